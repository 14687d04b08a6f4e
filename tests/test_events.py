"""Event -> frame integration (reference examples/input_pipeline.py:142-219) and sowed densities
(examples/tcja/models.py:128-142): oracle pinned by hand-computed cases and properties on the CPU,
CUDA kernels bit-exact against the oracle on the GPU (through the C-ABI)."""
import numpy as np
import pytest
import torch

from oracle import ref_events


def _events(rng, n, wh, p_out=0.0):
  a = np.stack([rng.integers(0, wh, n), rng.integers(0, wh, n), rng.integers(0, 2, n)], 1).astype(np.int32)
  if p_out and n:
    k = rng.random(n) < p_out
    a[k, 0] = wh + rng.integers(0, 5, k.sum())          # outside the sensor
  return a


def test_oracle_hand_case_split_by_number():
  # 7 events, T = 3 -> di = 2: frames take events [0,2), [2,4), [4,7)  (input_pipeline.py:166-178)
  addrs = np.array([[0, 0, 1], [1, 0, 0], [1, 1, 1], [1, 1, 1], [0, 1, 0], [0, 1, 0], [3, 3, 5]], np.int32)
  f = ref_events.preprocess_data_number(addrs, 3, 4)
  assert f.shape == (3, 4, 4, 2) and f.dtype == np.int32
  want = np.zeros((3, 4, 4, 2), np.int32)
  want[0, 0, 0, 1] = 1; want[0, 0, 1, 0] = 1            # [t, y, x, polarity != 0]
  want[1, 1, 1, 1] = 2
  want[2, 1, 0, 0] = 2; want[2, 3, 3, 1] = 1            # p = 5 counts as "not 0"
  assert np.array_equal(f, want)
  # resolution_scale 2: (x, y) -> (x // 2, y // 2) on a 2 x 2 grid
  f2 = ref_events.preprocess_data_number(addrs, 3, 4, 2)
  assert f2.shape == (3, 2, 2, 2) and f2.sum() == 7 and f2[2, 1, 1, 1] == 1 and f2[1, 0, 0, 1] == 2


def test_oracle_properties():
  rng = np.random.default_rng(0)
  for n, T, wh in [(1000, 20, 16), (19, 20, 8), (0, 5, 8), (4001, 10, 32)]:
    a = _events(rng, n, wh)
    f = ref_events.preprocess_data_number(a, T, wh)
    assert f.sum() == n                                           # every in-range event lands once
    di = n // T
    per = f.reshape(T, -1).sum(1)
    assert list(per[:-1]) == [di] * (T - 1) and per[-1] == n - di * (T - 1)
    assert f[..., 1].sum() == (a[:, 2] != 0).sum()
  d = ref_events.sow_densities(np.array([[[1, 0, 0, 2]], [[0, 0, 0, 0]]]))     # (T=2, B=1, 4)
  assert d["min"] == 0.5 and d["mean"] == 0.25 and d["counts"].tolist() == [[2], [0]]


@pytest.mark.gpu
@pytest.mark.parametrize("wh,rs,T", [(128, 1, 20), (128, 4, 10), (64, 2, 7)])
def test_events_to_frames_bit_exact(cuda_lib, wh, rs, T):
  from snnquantprune_b200 import input_pipeline as ip
  rng = np.random.default_rng(wh + rs)
  sizes = [30000, 0, 13, T, 50001, 1]
  samples = [_events(rng, n, wh, p_out=0.01) for n in sizes]
  samples[4][:4000, :2] = [5, 9]                                   # one hot pixel -> uint8 saturation
  addrs, off = ip.concat_events(samples)
  want32 = ref_events.batch_to_frames(samples, T, wh, rs)
  want8, nsat = ref_events.batch_to_frames(samples, T, wh, rs, saturate_u8=True)
  for bound in (0, max(sizes), 10**9):          # 32-bit counters / 16-bit packed counters / bound too large to pack
    got32, _ = ip.events_to_frames(addrs.cuda(), off.cuda(), T, wh, rs, exact_int32=True, max_events_per_sample=bound)
    assert np.array_equal(got32.cpu().numpy(), want32), bound
    got8, sat = ip.events_to_frames(addrs.cuda(), off.cuda(), T, wh, rs, max_events_per_sample=bound)
    assert got8.dtype == torch.uint8 and np.array_equal(got8.cpu().numpy(), want8), bound
    assert int(sat.item()) == nsat and nsat > 0


@pytest.mark.gpu
def test_preprocess_data_number_reference_signature_and_errors(cuda_lib):
  from types import SimpleNamespace
  from snnquantprune_b200 import input_pipeline as ip
  rng = np.random.default_rng(5)
  a = _events(rng, 5000, 128)
  cfg = SimpleNamespace(num_frames=20, resolution_scale=1, split_by="number")
  f = ip.preprocess_data_number(a, None, cfg, 128)
  assert np.array_equal(f.cpu().numpy(), np.minimum(ref_events.preprocess_data_number(a, 20, 128), 255))
  with pytest.raises(NotImplementedError):
    ip.preprocess_data_number(a, None, SimpleNamespace(num_frames=20, split_by="time"), 128)
  with pytest.raises(ValueError):
    ip.events_to_frames(torch.zeros((4, 3), dtype=torch.int64).cuda(), torch.zeros(2, dtype=torch.int64).cuda(), 4, 32)
  from snnquantprune_b200 import _lib
  rc = _lib.lib().snnqp_events_to_frames(None, None, 1, 1, 32, 1, 0, None, 0, None, None)
  assert rc != 0 and b"null pointer" in _lib.lib().snnqp_last_error()


@pytest.mark.gpu
def test_events_feed_the_hot_path(cuda_lib):
  """events -> frames (GPU) -> forward == forward on the oracle-integrated frames."""
  from snnquantprune_b200 import CextNetEngine, pack_cextnet, synthetic, input_pipeline as ip
  bits, T, H, B = 8, 4, 32, 3
  rng = np.random.default_rng(9)
  samples = [_events(rng, n, H) for n in (900, 1500, 400)]
  addrs, off = ip.concat_events(samples)
  fr, _ = ip.events_to_frames(addrs.cuda(), off.cuda(), T, H)
  want = ref_events.batch_to_frames(samples, T, H, 1, saturate_u8=True)[0]
  v = synthetic.make_variables(bits=bits, prune_percentage=0.5, T=T, H=H, seed=1)
  eng = CextNetEngine(pack_cextnet(v, bits, T, H, device="cuda"), lif_mode=0)
  assert np.array_equal(eng.forward(fr).cpu().numpy(), eng.forward(torch.as_tensor(want).cuda()).cpu().numpy())


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(4, 3, 16, 16, 128), (2, 5, 7, 16), (20, 2, 64, 64, 128), (3, 4, 110), (2, 2, 5), (70001, 1, 48)])
def test_density_stats_bit_exact(cuda_lib, shape):
  from snnquantprune_b200 import input_pipeline as ip
  rng = np.random.default_rng(len(shape))
  x = (rng.random(shape) < 0.13).astype(np.uint8) * rng.integers(1, 256, shape).astype(np.uint8)
  d = ip.density_stats(torch.as_tensor(x).cuda(), shape[0] * shape[1])
  w = ref_events.sow_densities(x)
  assert np.array_equal(d["counts"].cpu().numpy().reshape(shape[:2]), w["counts"])
  assert abs(d["min"].item() - w["min"]) < 1e-12 and abs(d["mean"].item() - w["mean"]) < 1e-12


@pytest.mark.gpu
def test_engine_densities_match_oracle(cuda_lib):
  """CextNetEngine.densities() == the reference's sowed input densities computed by the oracle on the same tensors."""
  from snnquantprune_b200 import CextNetEngine, pack_cextnet, synthetic
  bits, T, H, B = 8, 4, 128, 3          # the reference geometry: bit-packed spikes between conv1 .. conv4
  v = synthetic.make_variables(bits=bits, prune_percentage=0.5, T=T, H=H, seed=1)
  fr = torch.as_tensor(synthetic.make_frames(B, T, H, H, seed=2)).cuda()
  eng = CextNetEngine(pack_cextnet(v, bits, T, H, device="cuda"), lif_mode=0)
  c = {}
  eng.forward(fr, collect=c)
  eng.forward(fr)                      # production pass: densities() reads ITS buffers (bit-packed s1 .. s3)
  d = eng.densities(fr)
  # names as the reference sows them (examples/tcja/models.py:128-173): block 4's input is conv_t_0_inpt
  for key, name in (("conv_0_inpt", None), ("conv_1_inpt", "s1"), ("conv_2_inpt", "s2"), ("conv_t_0_inpt", "s3"),
                    ("conv_t_1_inpt", "p4"), ("dense2_inpt", "d1")):
    x = fr.cpu().numpy() if name is None else c[name].cpu().numpy()
    w = ref_events.sow_densities(np.swapaxes(x, 0, 1))             # (T, B, ...)
    assert abs(d[key]["min"] - w["min"]) < 1e-12 and abs(d[key]["mean"] - w["mean"]) < 1e-12, key
  assert 0.0 < d["conv_1_inpt"]["mean"] < 1.0
  # the same densities counted by the producing epilogues (snnqp_block_params.y_popcount: popcount of the ballot words)
  if eng.packed_spikes:
    eng2 = CextNetEngine(pack_cextnet(v, bits, T, H, device="cuda"), lif_mode=0, track_densities=True)
    eng2.forward(fr)
    d2 = eng2.densities(fr)
    for key in ("conv_1_inpt", "conv_2_inpt", "conv_t_0_inpt"):
      assert abs(d2[key]["min"] - d[key]["min"]) < 1e-12 and abs(d2[key]["mean"] - d[key]["mean"]) < 1e-12, key


# ---- committed golden fixture (tests/golden/make_golden.py: events_fixture) ---------------------------------------
import os
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "events_T6_wh32.npz")


def _gold_samples(z):
  off = z["offsets"]
  return [z["addrs"][off[i]:off[i + 1]] for i in range(len(off) - 1)]


def test_oracle_matches_events_golden_fixture():
  z = np.load(GOLD)
  wh, T = int(z["wh"]), int(z["T"])
  samples = _gold_samples(z)
  for rs in (1, 2):
    assert np.array_equal(ref_events.batch_to_frames(samples, T, wh, rs), z[f"frames_i32_rs{rs}"])
    f8, nsat = ref_events.batch_to_frames(samples, T, wh, rs, saturate_u8=True)
    assert nsat == int(z[f"n_saturated_rs{rs}"]) and nsat > 0
    assert np.array_equal(ref_events.sow_densities(np.swapaxes(f8, 0, 1))["counts"], z[f"density_counts_rs{rs}"])
  # structure of the fixture: the empty sample gives empty frames, the 4-event sample puts everything in the last frame
  f = z["frames_i32_rs1"]
  assert f[1].sum() == 0 and f[2, :-1].sum() == 0 and f[2, -1].sum() == 4


@pytest.mark.gpu
def test_events_and_density_kernels_match_golden_fixture(cuda_lib):
  from snnquantprune_b200 import input_pipeline as ip
  z = np.load(GOLD)
  wh, T = int(z["wh"]), int(z["T"])
  samples = _gold_samples(z)
  addrs, off = ip.concat_events(samples)
  assert np.array_equal(addrs.numpy(), z["addrs"]) and np.array_equal(off.numpy(), z["offsets"])
  for rs in (1, 2):
    for bound in (0, 4000):
      f32, _ = ip.events_to_frames(addrs.cuda(), off.cuda(), T, wh, rs, exact_int32=True, max_events_per_sample=bound)
      assert np.array_equal(f32.cpu().numpy(), z[f"frames_i32_rs{rs}"])
      f8, sat = ip.events_to_frames(addrs.cuda(), off.cuda(), T, wh, rs, max_events_per_sample=bound)
      assert np.array_equal(f8.cpu().numpy(), np.minimum(z[f"frames_i32_rs{rs}"], 255).astype(np.uint8))
      assert int(sat.item()) == int(z[f"n_saturated_rs{rs}"])
    d = ip.density_stats(f8, f8.shape[0] * T)
    assert np.array_equal(d["counts"].cpu().numpy().reshape(f8.shape[0], T).T, z[f"density_counts_rs{rs}"])


# ---- zero-suppressed frames: the host -> device wire format of the end-to-end path --------------------------------
def _zsf_cases():
  from snnquantprune_b200 import synthetic
  fr = synthetic.make_frames(6, 4, 32, 32, seed=3)
  fr[1] = 0                                   # an empty sample
  fr[2] = np.maximum(fr[2], 1)                # every cell non-zero
  fr[3, 0, 0, 0, 0] = 15                      # the largest 4-bit count
  yield "nibble", fr
  fr = fr.copy()
  fr[4, 1, 2, 3, 1] = 255                     # forces 8-bit values
  yield "byte", fr
  yield "all_zero", np.zeros((2, 4, 32, 32, 2), np.uint8)


def test_zsf_round_trip_cpu():
  """encode (product, numpy) -> decode (oracle, numpy) is the identity for every sample range."""
  from snnquantprune_b200 import input_pipeline as ip2
  for name, fr in _zsf_cases():
    z = ip2.zsf_encode(fr)
    assert z.value_bits == (8 if name == "byte" else 4)
    B = fr.shape[0]
    for b0, b1 in ((0, B), (1, B - 1), (B - 1, B), (0, 1)):
      if b1 <= b0:
        continue
      bm, bo, v, base, nb = z.chunk(b0, b1)
      d = ref_events.zsf_decode(bm.numpy().view(np.uint32), bo.numpy().view(np.uint32), v.numpy(), base, nb, z.value_bits)
      assert np.array_equal(d, fr[b0:b1].reshape(-1)), (name, b0, b1)
    assert z.nbytes < fr.nbytes or name == "nibble" or name == "byte"
  with pytest.raises(ValueError):
    ip2.zsf_encode(np.zeros((1, 3, 5, 5, 2), np.uint8))       # cells not a multiple of 1024


@pytest.mark.gpu
def test_zsf_expand_kernel_bit_exact(cuda_lib):
  from snnquantprune_b200 import input_pipeline as ip2
  for name, fr in _zsf_cases():
    z = ip2.zsf_encode(fr)
    B = fr.shape[0]
    for b0, b1 in ((0, B), (1, B - 1), (B - 1, B)):
      if b1 <= b0:
        continue
      bm, bo, v, base, nb = z.chunk(b0, b1)
      out = torch.full((nb * 1024,), 7, device="cuda", dtype=torch.uint8)
      vd = torch.cat([v, torch.zeros(16, dtype=torch.uint8)]).cuda()
      ip2.zsf_expand(bm.cuda(), bo.cuda(), vd, base, nb, z.value_bits, out)
      assert np.array_equal(out.cpu().numpy(), fr[b0:b1].reshape(-1)), (name, b0, b1)


@pytest.mark.gpu
def test_forward_host_zsf_equals_device_forward(cuda_lib):
  """The end-to-end host call in the zero-suppressed wire format gives exactly the logits of the device-resident
  forward and of the dense-frame host call (reference geometry, ragged chunk schedule)."""
  from snnquantprune_b200 import CextNetEngine, pack_cextnet, synthetic
  from snnquantprune_b200 import input_pipeline as ip2
  bits, T, H, B = 8, 20, 128, 45
  v = synthetic.make_variables(bits=bits, prune_percentage=0.5, T=T, H=H, seed=1)
  fr = synthetic.make_frames(B, T, H, H, seed=5)
  eng = CextNetEngine(pack_cextnet(v, bits, T, H, device="cuda"), chunk=37)
  want = eng.forward(torch.as_tensor(fr).cuda()).cpu()
  z = ip2.zsf_encode(fr)
  assert z.nbytes < 0.3 * fr.nbytes
  host_out = torch.empty_like(want).pin_memory()
  eng.forward_host_zsf(z, host_out)
  torch.cuda.synchronize()
  assert torch.equal(host_out, want)
  eng.forward_host(torch.as_tensor(fr).pin_memory(), host_out)
  torch.cuda.synchronize()
  assert torch.equal(host_out, want)
