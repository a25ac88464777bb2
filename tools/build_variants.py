"""Occupancy experiment builds of librbphd.so (launch shape overrides, see rbphd_block.cuh).

    python tools/build_variants.py [name ...]
"""
import os
import sys
from concurrent.futures import ThreadPoolExecutor

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monorfs_b200 import build as B  # noqa: E402

# RBPHD_GRID_CELLS must stay >= 4096: the exploration mini-grid of phase A4 asks for up to 4096 cells of the shared
# offset buffer (rbphd_kernels.cu, grid_build(..., 4096))
HALF = dict(RBPHD_SORT_CAP=4096, RBPHD_VS_CAP=2048, RBPHD_SORT_BUCKETS=2048, RBPHD_GRID_CELLS=4096)
VARIANTS = {
    "b1024x1": dict(RBPHD_BLOCK=1024, RBPHD_CTAS_PER_SM=1),
    "b512x1": dict(RBPHD_BLOCK=512, RBPHD_CTAS_PER_SM=1),
    "b512x2": dict(RBPHD_BLOCK=512, RBPHD_CTAS_PER_SM=2, **HALF),
    "b384x2": dict(RBPHD_BLOCK=384, RBPHD_CTAS_PER_SM=2, **HALF),
    "b256x2": dict(RBPHD_BLOCK=256, RBPHD_CTAS_PER_SM=2, **HALF),
}

if __name__ == "__main__":
    names = sys.argv[1:] or list(VARIANTS)
    with ThreadPoolExecutor(max_workers=4) as ex:
        for name, lib in zip(names, ex.map(lambda n: B.build_variant(n, VARIANTS[n]), names)):
            print(name, lib)
