"""cudasbmp_b200 — the KGMT tree-expansion hot path of nipe1783/cudaSBMP, written from scratch for
NVIDIA B200 (sm_100a).  Product = cudasbmp_b200/libkgmt_b200.so (C ABI: include/kgmt_c.h) built from
cudasbmp_b200/csrc; this package is its Python face.  Nothing here imports oracle/."""
from .kgmt import KGMT, KgmtError, load, LIB_PATH, ABI_SYMBOLS  # noqa: F401
from . import workloads  # noqa: F401
