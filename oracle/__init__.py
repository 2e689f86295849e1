"""
oracle/ -- CPU restatement of the reference (Bismarrck/tensoralloy) energy /
force / virial / Hessian hot path.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import it.  Nothing under `tensoralloy_b200/` imports it and the
product path raises if its CUDA library is missing -- there is no CPU fallback.

What it is: numpy (integer / index work, neighbour lists) and torch-float64 on
the CPU (floating-point model arithmetic; `torch.autograd` stands in for the
`tf.gradients` / `tf.hessians` calls of the reference, `nn/basic.py:281,306,418`).
Every function cites the reference file:line it restates.

Pinning status (see DESIGN.md section "Oracle"):
  * zjw04 rho/phi/F       pinned to test_files/lammps/zjw04_Ni.alloy.eam and
                          Zhou_AlCu.alloy.eam (tables written by the reference)
  * G2/G4 descriptors     pinned to test_files/amp_Pd3O2.npz
  * Hessian               pinned to test_files/crystals/Ni_fc2.npy (fp32 golden)
  * elastic constants     pinned to the values asserted in
                          nn/constraint/tests/test_elastic.py:49-56
  * neighbour lists       third-party (ASE >= 3.21 `neighbor_list`, absent from
                          /root/reference): published algorithm restated, parity
                          anchored on tests/test_neighbor.py-style counts and on
                          the goldens above (a wrong list changes E/F/H).
  * E/F/stress vs LAMMPS  UNPINNED here (no LAMMPS in the container).
The goldens are copied as small fixtures to tests/golden/ by
tests/golden/make_golden.py (the reference tree is not available on the GPU box).
"""
