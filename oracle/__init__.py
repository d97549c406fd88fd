"""CPU oracle for the MHLA + SPPP hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU, the algorithm of the reference
(zser092/Focused-Attention-ViT, `models/mhla.py` and `models/sppp.py`).  It is the
checker that the CUDA path is compared against.  Only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import it.  The product package (`focused-attention-vit_b200/`,
imported as `favit_b200`) never imports anything from here and fails loudly when
its CUDA library is missing.

Parity pin: the reference ships no tests or golden vectors, so the
oracle is pinned against outputs of the reference itself, executed in the build
container by `tests/golden/make_golden.py`; the resulting fixtures live in
`tests/golden/*.npz` and `tests/test_oracle_golden.py` checks the oracle against
them on every CPU run.
"""
from .mhla_oracle import (  # noqa: F401
    window_indices,
    window_multiplicity,
    mhla_forward_gather,
    mhla_forward_closed_form,
    mhla_attn_core_closed_form,
    mhla_attn_core_gather,
    dropout_keep_mask,
    mlp_dropout_keep_mask,
    fold_latent,
)
from .models_oracle import (  # noqa: F401
    vit_mhla_forward,
    sppp_vit_mhla_forward,
    superpixel_centroids,
    dynamic_positional_encoding,
)
from .slic_oracle import slic_oracle, slic_grid  # noqa: F401
from .sppp_oracle import (  # noqa: F401
    map_patches_oracle,
    assign_oracle,
    pool_mean_oracle,
    pool_mean_batched_oracle,
    pool_variant_batched_oracle,
)
