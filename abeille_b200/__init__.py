"""abeille_b200 -- B200-native transport backend for the Abeille Monte Carlo code.

The product is two native libraries built in-tree (``python -c 'import __graft_entry__ as g; g.build()'``):

* ``lib/libabeille_b200.so``  hand-written sm_100a CUDA kernels behind the C ABI of ``include/abeille_b200.h``
* ``lib/libabeille_host.so``  the C++ host: YAML deck -> object model -> GPUTransporter / PowerIterator

This package is only the ctypes plumbing over those two libraries (plus torch for device memory and
torch.distributed in :mod:`abeille_b200.distributed`).  There is no CPU fallback: every compute call ends
in the CUDA library and raises if it is missing or no device is present.
"""
from .backend import (Backend, BackendError, lib_paths, load_backend_lib, load_host_lib, new_bank, parse_only,
                      dump_tables, source_records, yaml_roundtrip, comb_particles, comb_rows, global_rng_state, BANK_F64, BANK_U64, COUNTER_KEYS)

__all__ = ["Backend", "BackendError", "lib_paths", "load_backend_lib", "load_host_lib", "new_bank", "parse_only",
           "dump_tables", "source_records", "yaml_roundtrip", "comb_particles", "comb_rows", "global_rng_state", "BANK_F64", "BANK_U64", "COUNTER_KEYS"]
