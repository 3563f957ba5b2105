"""flow-pipeline_b200 -- B200-native flow-aggregation stage (drop-in for the
decode->aggregate path of cloudflare/flow-pipeline).

The package holds only what that path needs:
  csrc/         the sm_100a CUDA kernels and the C ABI (include/flowagg.h)
  flowagg.py    ctypes binding of libflowagg.so (what the tests and bench.py drive)
  parallel.py   one-process-per-GPU plumbing over torch.distributed (sketch all-reduce, row exchange)
  host/         inserter.cc: C++ mirror of inserter/inserter.go's consumer-group handler over the C ABI;
                go/inserter_b200.go: the cgo shim (source only, no Go toolchain in this image)

The directory name contains a hyphen (it mirrors the reference's name), so
import it with importlib.import_module("flow-pipeline_b200") or through the
flow_pipeline_b200 shim module at the repository root.
"""
from .flowagg import (  # noqa: F401
    FaConfig, FaMockerConfig, FlowAgg, FlowAggError, KEY_MODES, KEY_WORDS, ROW_DTYPE, HH_DTYPE,
    build, lib_path, load_library, mocker_host, row_owner,
    FA_ADDR_MOCKER, FA_ADDR_ZIPF24, FA_ADDR_UNIQUE, FA_CMS_LOCAL, FA_CMS_GLOBAL,
)
