"""ctypes binding of the CPU oracle (oracle/libpedfac_oracle.so). TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
It reuses the product's generic C-ABI wrapper class (the oracle exports the same entry points
with a `pfo_` prefix); the product never imports this module."""
import ctypes as C
from pathlib import Path

from pedfac_b200.abi import CEngine, bind

_lib = {}


def load_oracle_library(opt: str = "O2") -> C.CDLL:
    if opt not in _lib:
        name = "libpedfac_oracle.so" if opt == "O2" else "libpedfac_oracle_O0.so"
        path = Path(__file__).resolve().parent / name
        if not path.exists():
            raise FileNotFoundError(f"{path} not found: run `make -C oracle`")
        lib = C.CDLL(str(path))
        bind(lib, "pfo_")
        lib.pfo_tprob.restype = C.POINTER(C.c_double)
        _lib[opt] = lib
    return _lib[opt]


class OracleEngine(CEngine):
    def __init__(self, n_loci, cap_indiv, cap_marr, locus_range=None, n_chains=1, device=0, opt="O2"):
        assert n_chains == 1, "the oracle is single-chain; use one OracleEngine per chain"
        super().__init__(load_oracle_library(opt), "pfo_", n_loci, cap_indiv, cap_marr, locus_range, 1, 0)


def tprob():
    p = load_oracle_library().pfo_tprob()
    return [p[i] for i in range(27)]
