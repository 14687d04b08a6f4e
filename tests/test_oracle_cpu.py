"""CPU tests of the oracle itself: the restatements against each other, against
the committed golden vectors, and against the properties the reference's own
tests pin for neighbouring code (quant_test.py:146-250 integer-grid round trip
and level count; flax_qconv_test.py:148-285 conv geometry == a trusted conv)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import ref_int, ref_net, ref_quant, ref_snn
from snnquantprune_b200 import synthetic

GOLD = os.path.join(os.path.dirname(__file__), "golden")
F32 = np.float32


def test_duq_golden_vectors(oracle_lib):
  g = json.load(open(os.path.join(GOLD, "duq_vectors.json")))
  w = np.array(g["w"], F32); mask = np.array(g["mask"], F32)
  for case in g["cases"]:
    q = ref_quant.duq_levels(w, case["a"], case["bits"])
    assert q.tolist() == case["levels"]
    qc = ref_int.duq_levels_c(w, None, case["a"], case["bits"])
    assert qc.astype(np.int32).tolist() == case["levels"]          # C == numpy restatement
    fw = ref_quant.effective_weight(w, case["a"], case["c"], mask, case["bits"])
    assert np.array_equal(fw, np.array(case["forward"], F32))


@pytest.mark.parametrize("bits", [2, 3, 4, 6, 8])
def test_duq_level_count_and_symmetry(bits):
  # restates quant_test.py:193-250 for DuQ: at most 2^bits - 1 distinct levels,
  # symmetric grid, zero preserved, values beyond +-a saturate.
  rng = np.random.default_rng(bits)
  w = rng.standard_normal(20000).astype(F32)
  a = F32(1.3)
  q = ref_quant.duq_levels(w, a, bits)
  L = 2 ** (bits - 1) - 1
  assert q.min() == -L and q.max() == L
  assert len(np.unique(q)) <= 2 ** bits - 1
  assert np.array_equal(ref_quant.duq_levels(-w, a, bits), -q)
  assert ref_quant.duq_levels(np.zeros(3, F32), a, bits).tolist() == [0, 0, 0]
  fw = ref_quant.duq_forward(w, a, a, bits)
  assert np.max(np.abs(fw)) <= a * (1 + 1e-6)


def test_duq_integer_grid_round_trip():
  # restates quant_test.py:146-185: data already on the integer grid survives.
  for bits in (4, 8):
    L = 2 ** (bits - 1) - 1
    ints = np.arange(-L, L + 1).astype(F32)
    a = F32(L)
    assert np.array_equal(ref_quant.duq_levels(ints, a, bits), ints.astype(np.int32))
    assert np.array_equal(ref_quant.duq_forward(ints, a, a, bits), ints)


def test_duq_round_half_even_and_passthrough():
  # L = 1 (2 bits): x * L = +-0.5 must round to 0 (half to even), 1.5 -> 2 is clipped
  assert ref_quant.duq_levels(np.array([0.5, -0.5, 0.5000001], F32), 1.0, 2).tolist() == [0, 0, 1]
  w = np.array([0.3, -2.0], F32)
  assert np.array_equal(ref_quant.duq_forward(w, -1.0, -1.0, 8), w)        # a == -1
  assert np.array_equal(ref_quant.duq_forward(w, 1.0, 1.0, -1), w)          # bits == -1


def test_effective_weight_matches_sparsity_py_formula():
  rng = np.random.default_rng(3)
  k = (rng.standard_normal((3, 3, 8, 16)) * 0.2).astype(F32)
  a = ref_quant.gaussian_init(k, 8)
  m = ref_quant.local_mask(k, 0.5)
  assert np.array_equal(ref_quant.effective_weight(k, a, a, m, 8),
                        ref_quant.effective_weight_sparsity_py(k, a, m, 8))


def test_masks_local_and_global():
  rng = np.random.default_rng(4)
  ks = {"QuantConv_0": rng.standard_normal((3, 3, 2, 8)).astype(F32),
        "QuantConv_1": (rng.standard_normal((3, 3, 8, 8)) * 0.1).astype(F32),
        "QuantDense_0": rng.standard_normal((16, 4)).astype(F32)}
  m = ref_quant.local_mask(ks["QuantConv_1"], 0.3)
  k = int(ks["QuantConv_1"].size * 0.3)
  assert int((m == 0).sum()) == k
  assert np.abs(ks["QuantConv_1"])[m == 0].max() <= np.abs(ks["QuantConv_1"])[m == 1].min()
  gm = ref_quant.global_masks(ks, 0.5)
  total = sum(v.size for v in ks.values())
  assert sum(int((v == 0).sum()) for v in gm.values()) == int(total * 0.5)
  thr = max(np.abs(ks[n])[gm[n] == 0].max() for n in ks if (gm[n] == 0).any())
  assert all(np.abs(ks[n])[gm[n] == 1].min() >= thr for n in ks)
  # the small-magnitude layer loses the most weights under a global ranking
  assert (gm["QuantConv_1"] == 0).mean() > (gm["QuantConv_0"] == 0).mean()
  # product-side host mirror gives the same masks
  from snnquantprune_b200 import quant as hq
  gm2 = hq.global_masks(ks, 0.5)
  assert all(np.array_equal(gm[n], gm2[n]) for n in ks)
  assert hq.gaussian_init(ks["QuantConv_1"], 8) == ref_quant.gaussian_init(ks["QuantConv_1"], 8)


def test_gaussian_init():
  x = np.array([0.0, 0.0, 0.0], F32)
  assert ref_quant.gaussian_init(x, 4) == F32(1 / 16)
  rng = np.random.default_rng(0)
  x = rng.standard_normal(1000).astype(F32)
  mu, sd = x.mean(), x.std()
  assert np.isclose(ref_quant.gaussian_init(x, 8), max(abs(mu - 3 * sd), abs(mu + 3 * sd)), rtol=1e-6)


@pytest.mark.parametrize("shape", [(1, 2, 2, 4, 8), (2, 6, 4, 2, 16), (1, 8, 8, 128, 32), (3, 4, 10, 12, 8)])
def test_conv3x3_geometry_matches_trusted_conv(oracle_lib, shape):
  # flax_qconv_test.py:148-285 restated: integer conv == torch conv2d (tol 0)
  N, H, W, Cin, Cout = shape
  rng = np.random.default_rng(N * 7 + H)
  x = rng.integers(0, 4, size=(N, H, W, Cin)).astype(np.uint8)
  q = rng.integers(-127, 128, size=(3, 3, Cin, Cout)).astype(np.int8)
  acc = ref_int.conv3x3_acc(x, q)
  xt = torch.from_numpy(x.astype(np.float64)).permute(0, 3, 1, 2)
  wt = torch.from_numpy(q.astype(np.float64)).permute(3, 2, 0, 1)
  ref = torch.nn.functional.conv2d(xt, wt, padding=1).permute(0, 2, 3, 1).numpy()
  assert np.array_equal(acc, ref.astype(np.int32))


def test_same_padding_k4_is_1_2_and_conv1d(oracle_lib):
  assert ref_snn.same_pads(20, 4) == (1, 2)
  assert ref_snn.same_pads(128, 4) == (1, 2)
  rng = np.random.default_rng(5)
  x = rng.integers(0, 50, size=(2, 9, 6)).astype(np.int32)
  q = rng.integers(-7, 8, size=(4, 6, 5)).astype(np.int8)
  acc = ref_int.conv1d_acc(x, q, 1)
  ref = ref_snn.conv1d_nwc(x.astype(F32), q.astype(F32), (1, 2))
  assert np.array_equal(acc, ref.astype(np.int32))


def test_lif_c_matches_numpy_and_threshold_is_inclusive(oracle_lib):
  rng = np.random.default_rng(6)
  T, M, C = 7, 33, 16
  acc = rng.integers(-300, 300, size=(T, M, C)).astype(np.int32)
  scale = rng.uniform(0.001, 0.01, C).astype(F32); bias = rng.uniform(-0.2, 0.8, C).astype(F32)
  s_c, u_c, pre = ref_int.lif_from_acc(acc, scale, bias, want_pre=True)
  v = ref_int.fmaf(acc.astype(F32), scale[None, None, :], bias[None, None, :])
  assert np.array_equal(pre, v)
  u = np.zeros((M, C), F32)
  for t in range(T):
    u, s = ref_snn.lif_step(u, v[t])
    assert np.array_equal(s.astype(np.uint8), s_c[t])
  assert np.array_equal(u, u_c)
  # u lands exactly on the threshold: (0 + (2 - 0)/2) = 1 -> spike, reset to 0
  s1, u1 = ref_int.lif_from_acc(np.array([[[2]]], np.int32), np.ones(1, F32), np.zeros(1, F32))
  assert s1.item() == 1 and u1.item() == 0.0
  below = np.nextafter(F32(2), F32(0))
  s2, _ = ref_int.lif_from_acc(np.array([[[below]]], F32), np.ones(1, F32), np.zeros(1, F32))
  assert s2.item() == 0


def test_fmaf_is_single_rounding(oracle_lib):
  rng = np.random.default_rng(7)
  a = rng.standard_normal(5000).astype(F32) * 1000
  b = rng.standard_normal(5000).astype(F32)
  c = rng.standard_normal(5000).astype(F32)
  got = ref_int.fmaf(a, b, c)
  from fractions import Fraction
  for i in range(0, 5000, 250):                 # exact rational check on a sample
    exact = Fraction(float(a[i])) * Fraction(float(b[i])) + Fraction(float(c[i]))
    cand = F32(float(exact))                   # float(Fraction) rounds once to double
    assert abs(Fraction(float(got[i])) - exact) <= abs(Fraction(float(cand)) - exact)


def test_maxpool_and_vote(oracle_lib):
  rng = np.random.default_rng(8)
  s = (rng.uniform(size=(2, 3, 4, 6, 8)) < 0.3).astype(np.uint8)
  assert np.array_equal(ref_int.maxpool2_u8(s), ref_snn.maxpool2(s.astype(F32)).astype(np.uint8))
  sp = (rng.uniform(size=(20, 4, 110)) < 0.2).astype(np.uint8)
  v = ref_int.vote(sp)
  x = np.mean(sp.astype(F32), 0)
  ref = np.mean(x.reshape(4, 11, 10), -1)
  assert np.allclose(v, ref, rtol=0, atol=1e-7)


@pytest.mark.parametrize("cfg", [(8, 0.5, 5, 32, 2), (4, 0.8, 3, 64, 1), (2, 0.9, 3, 32, 2)])
def test_int_path_agrees_with_float_path(oracle_lib, cfg):
  """The two independent restatements: integer accumulators + folded affine vs
  the reference's fp32 op order.  North-star tolerances: membrane <= 1e-5
  relative, spike flip rate <= 1e-4."""
  bits, p, T, H, B = cfg
  v = synthetic.make_variables(bits=bits, prune_percentage=p, T=T, H=H, seed=21)
  fr = synthetic.make_frames(B, T, H, H, seed=22)
  pk = ref_net.pack_network(v, bits, H)
  ci, cf = {}, {}
  li = ref_net.forward(pk, fr, collect=ci)
  lf = ref_snn.cextnet_forward(v, fr, bits, collect=cf)
  for a, b in (("s1", "pool1"), ("s2", "pool2"), ("s3", "pool3"), ("s4", "conv4_spikes"),
               ("s5", "conv5_spikes"), ("d1", "dense1_spikes"), ("d2", "dense2_spikes")):
    assert np.mean(ci[a] != (cf[b] != 0)) <= 1e-4, a
    assert ci[a].mean() > 0.01, f"{a} is dead: parity would be vacuous"
  for i in range(1, 6):
    ui, uf = ci[f"conv{i}_u"], cf[f"conv{i}_u"]
    rel = np.abs(ui - uf) / np.maximum(np.abs(uf), 1.0)
    assert np.quantile(rel, 0.9999) <= 1e-5, (i, rel.max())
  assert np.abs(li - lf).max() <= 1e-6
  assert np.abs(ci["att4"] - cf["tcja1_att"]).max() <= 1e-5


@pytest.mark.parametrize("name", ["cextnet_T4_H32_b8_p50", "cextnet_T3_H32_b4_p80", "cextnet_T3_H32_b2_p90"])
def test_network_golden(oracle_lib, name):
  z = np.load(os.path.join(GOLD, name + ".npz"))
  m = json.loads(str(z["meta"]))
  v = synthetic.make_variables(bits=m["bits"], prune_percentage=m["prune"], T=m["T"], H=m["H"], seed=m["seed_w"])
  fr = synthetic.make_frames(m["B"], m["T"], m["H"], m["H"], seed=m["seed_x"])
  import sys
  sys.path.insert(0, GOLD)
  import make_golden
  if make_golden.variables_digest(v) != m["variables_sha"] or make_golden.sha(fr) != m["frames_sha"]:
    pytest.skip("numpy RNG stream differs from the one the fixture was made with")
  c = {}
  logits = ref_net.forward(ref_net.pack_network(v, m["bits"], m["H"]), fr, collect=c)
  assert np.array_equal(logits, z["logits_int"])
  for k in ("s1", "s2", "s3", "s4", "s5", "d1", "d2"):
    assert np.array_equal(np.packbits(c[k].reshape(-1)), z[k + "_bits"]), k
  assert np.array_equal(c["att4"], z["att4"])
  a = c["conv2_acc"].astype(np.int64)
  assert [a.sum(), (a ** 2).sum()] == z["conv2_acc_sum"].tolist()
