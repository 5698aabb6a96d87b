"""pedfac_b200 — B200-native likelihood engine for the pedFac `pedigraph` sampler's hot path.

The product is `pedfac_b200/lib/libpedfac_cuda.so` (hand-written sm_100a CUDA behind the C ABI
of `include/pedfac_cuda.h`). This package is the thin host-side mirror used by the tests and
`bench.py`: ctypes marshalling (`abi`), the pedigree graph bookkeeping that feeds the engine
(`pedigree`) and the synthetic workload generator (`synth`). There is no CPU fallback.
"""
from .abi import (CEngine, GraphView, PedfacError, PROPOSAL_DTYPE, PF_PROP_PRONG, PF_PROP_SELFING,
                  PF_PROP_TRIO, compose_ll, load_cuda_library, make_proposals)
from .pedigree import Marriage, Pedigree


class Engine(CEngine):
    """One libpedfac_cuda engine (one GPU, one locus shard). Raises if the CUDA library or a GPU
    is missing — by design."""

    def __init__(self, n_loci, cap_indiv, cap_marr, locus_range=None, n_chains=1, device=0):
        super().__init__(load_cuda_library(), "pf_", n_loci, cap_indiv, cap_marr, locus_range, n_chains, device)


__all__ = ["Engine", "CEngine", "GraphView", "PedfacError", "PROPOSAL_DTYPE", "PF_PROP_TRIO",
           "PF_PROP_SELFING", "PF_PROP_PRONG", "compose_ll", "load_cuda_library", "make_proposals",
           "Pedigree", "Marriage"]
