// Activation functions of the per-atom networks: value and derivative
// (reference nn/utils.py:20-60, tf.nn.*; ids = TAB_ACT_* of include/tab200.h).
#pragma once

template <typename Real>
__device__ __forceinline__ Real act_fn(int kind, Real z, Real &d) {
    switch (kind) {
    case 0: {   // softplus
        const Real e = exp(-fabs(z));
        const Real sp = (z > Real(0) ? z : Real(0)) + log1p(e);
        d = z >= Real(0) ? Real(1) / (Real(1) + e) : e / (Real(1) + e);
        return sp;
    }
    case 1: {   // tanh
        const Real t = tanh(z);
        d = Real(1) - t * t;
        return t;
    }
    case 2:     // relu
        d = z > Real(0) ? Real(1) : Real(0);
        return z > Real(0) ? z : Real(0);
    case 3:     // leaky_relu (tf default alpha 0.2)
        d = z > Real(0) ? Real(1) : Real(0.2);
        return z > Real(0) ? z : Real(0.2) * z;
    case 4: {   // sigmoid
        const Real s = Real(1) / (Real(1) + exp(-z));
        d = s * (Real(1) - s);
        return s;
    }
    case 5: {   // softsign
        const Real q = Real(1) + fabs(z);
        d = Real(1) / (q * q);
        return z / q;
    }
    case 6: {   // elu
        const Real e = exp(z);
        d = z > Real(0) ? Real(1) : e;
        return z > Real(0) ? z : e - Real(1);
    }
    default: {  // 7 squareplus (nn/utils.py:39-47)
        const Real s = sqrt(z * z + Real(4));
        d = Real(0.5) * (Real(1) + z / s);
        return Real(0.5) * (z + s);
    }
    }
}
