"""BatchUniversalTransformer (transformer/universal.py:921-1388) over the batched GPU lists."""
import numpy as np
import pytest

from tensoralloy_b200.atoms import Atoms, bulk_fcc
from tensoralloy_b200.transformer import UniversalTransformer

pytestmark = pytest.mark.gpu


def test_batch_universal_transformer_serves_batches_and_enforces_nij_max():
    """The reference's mini-batch transformer class as the transformer of a model: the same
    batch results as with UniversalTransformer; `nij_max` is enforced per structure."""
    from collections import Counter
    from tensoralloy_b200.transformer import BatchUniversalTransformer
    rng = np.random.default_rng(4)
    images = []
    for k in range(3):
        base = bulk_fcc('Ni', 3.52, (2, 2, 2 + k))
        images.append(Atoms(['Ni'] * len(base),
                            base.positions + rng.normal(scale=0.05, size=base.positions.shape),
                            base.cell, True))
    blf = BatchUniversalTransformer(Counter({'Ni': 64}), rcut=4.5, batch_size=8)
    feats = blf.get_batch_features(images)
    ref = UniversalTransformer(['Ni'], rcut=4.5).get_batch_features(images)
    c1, c2 = feats.nbr.counts().cpu().numpy(), ref.nbr.counts().cpu().numpy()
    assert np.array_equal(c1, c2) and c1.sum() > 0
    per_structure = [int(c1[feats.offsets[s]:feats.offsets[s + 1]].sum()) for s in range(3)]
    tight = BatchUniversalTransformer(Counter({'Ni': 64}), rcut=4.5, batch_size=8,
                                      nij_max=max(per_structure))
    tight.get_batch_features(images)
    short = BatchUniversalTransformer(Counter({'Ni': 64}), rcut=4.5, batch_size=8,
                                      nij_max=max(per_structure) - 1)
    with pytest.raises(ValueError, match="nij_max"):
        short.get_batch_features(images)
