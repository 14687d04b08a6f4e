"""int8 tcgen05.mma rate vs tile width N (M = 128, K = 32): SNNQP_PEAK_N=<N> python tools/imma_rate.py"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from snnquantprune_b200 import _lib
t = ctypes.c_double(0)
_lib.check(_lib.lib().snnqp_diag_imma_peak(4000, 3, ctypes.byref(t), _lib.stream()))
print(f"N={os.environ.get('SNNQP_PEAK_N', '256')}: {t.value:.1f} TOP/s")
