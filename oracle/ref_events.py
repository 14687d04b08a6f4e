"""ORACLE (test infrastructure only -- never imported by the product path).

CPU restatement of the reference's event -> frame integration, "split by
number" (/root/reference/examples/input_pipeline.py:142-219,
``preprocess_data_number``), and of the activation densities the model sows
(/root/reference/examples/tcja/models.py:128-142).  pinned to the executed reference: the
reference ships no test or fixture for either and cannot run here (no
tensorflow / jax); the restatement follows the source line by line.
"""
from __future__ import annotations

import numpy as np


def preprocess_data_number(addrs: np.ndarray, num_frames: int, wh: int, resolution_scale: float = 1) -> np.ndarray:
  """addrs: (N, 3) integer (x, y, p) in time order.  Returns int32 (T, wh', wh', 2),
  wh' = int(wh // resolution_scale)  (input_pipeline.py:157-158, 211-213)."""
  addrs = np.asarray(addrs)
  wh = int(wh // resolution_scale)                                           # :157
  frames = np.zeros((num_frames, 2, wh * wh), dtype=np.int32)                # :158
  n = addrs.shape[0]
  di = n // num_frames                                                       # :166
  j_l = np.array([i * di for i in range(num_frames)], dtype=np.int64)        # :167-169
  j_r = np.array([j_l[i] + di if i < num_frames - 1 else n for i in range(num_frames)], dtype=np.int64)  # :170-178
  for i in range(num_frames):
    x = addrs[j_l[i]: j_r[i], 0].astype(np.float32) // np.float32(resolution_scale)     # :183-186 (float floor-div)
    y = addrs[j_l[i]: j_r[i], 1].astype(np.float32) // np.float32(resolution_scale)     # :187-190
    p = addrs[j_l[i]: j_r[i], 2]                                                        # :191
    mask = [p == 0, np.logical_not(p == 0)]                                             # :192-194
    for j in range(2):
      position = (y[mask[j]] * wh + x[mask[j]]).astype(np.int32)                        # :196-199
      position = position[(position >= 0) & (position < wh * wh)]        # the reference's scatter would fail on these
      counts = np.bincount(position, minlength=0)                                       # :197-199
      frames[i, j, : counts.shape[0]] += counts.astype(np.int32)                        # :201-214 scatter_nd_add
  return np.transpose(frames.reshape(num_frames, 2, wh, wh), (0, 2, 3, 1))              # :216-218


def batch_to_frames(samples, num_frames: int, wh: int, resolution_scale: float = 1, saturate_u8: bool = False):
  """List of (N_b, 3) arrays -> (B, T, wh', wh', 2); optionally the product's uint8 saturation and its count."""
  out = np.stack([preprocess_data_number(a, num_frames, wh, resolution_scale) for a in samples])
  if not saturate_u8:
    return out
  return np.minimum(out, 255).astype(np.uint8), int((out > 255).sum())


def sow_densities(x: np.ndarray):
  """x: (T, B, ...) activations.  models.py:128-134: per (t, b) fraction of non-zeros, then ('_min' = max, '_mean')."""
  T, B = x.shape[:2]
  nz = (x.reshape(T, B, -1) != 0).sum(-1)
  frac = nz / np.prod(x.shape[2:])
  return {"counts": nz.astype(np.int64), "min": float(frac.max()), "mean": float(frac.mean())}


def zsf_decode(bitmap_u32: np.ndarray, block_off: np.ndarray, values: np.ndarray, value_base: int, n_blocks: int,
               value_bits: int) -> np.ndarray:
  """Decoder of the zero-suppressed frame wire format (include/snnqp.h, snnqp_expand_frames_zsf), plain numpy:
  returns the n_blocks * 1024 dense uint8 cells."""
  bits = np.unpackbits(np.ascontiguousarray(bitmap_u32[:n_blocks * 32]).view(np.uint8), bitorder="little").astype(bool)
  if value_bits == 4:
    v = np.stack([values & 0xF, values >> 4], -1).reshape(-1)
  else:
    v = values
  out = np.zeros(n_blocks * 1024, np.uint8)
  for k in range(n_blocks):
    m = bits[k * 1024:(k + 1) * 1024]
    start = int(block_off[k]) - int(value_base)
    out[k * 1024:(k + 1) * 1024][m] = v[start:start + int(m.sum())]
  return out
