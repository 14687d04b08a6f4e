"""The oracle (and the product's host-side numpy code) against fixtures made by
EXECUTING THE UNMODIFIED REFERENCE (tests/golden/make_from_reference.py, numpy
shims for jax / flax): this is what pins the oracle.  Elementwise pieces are
bit-exact; the whole network is held to the north-star tolerances against the
reference's own CextNet.__call__."""
import json
import os

import numpy as np
import pytest

import reffix
from oracle import ref_int, ref_net, ref_quant, ref_snn
from snnquantprune_b200 import quant as hq
from snnquantprune_b200.synthetic import StableRNG

F32 = np.float32
OPS = np.load(os.path.join(reffix.GOLD, "from_reference_ops.npz"))
META = json.loads(str(OPS["meta"]))


def test_duq_forward_bit_exact_vs_reference():
  w = OPS["duq_w"]
  n_half = 0
  for case in META["duq_cases"]:
    ref = OPS[case["key"]]
    a, c, bits = F32(case["a"]), F32(case["c"]), case["bits"]
    assert np.array_equal(ref_quant.duq_forward(w, a, c, bits).view(np.uint32), ref.view(np.uint32)), case
    # the integer levels the pack step stores reproduce the reference's quantized weights exactly
    L = 2 ** (bits - 1) - 1
    q = ref_quant.duq_levels(w, a, bits)
    assert np.array_equal(((q.astype(F32) / F32(L)) * c).astype(F32), ref), case
    x = np.clip(w.astype(np.float64) / np.float64(a), -1, 1) * L
    n_half += int(np.sum(np.abs(np.abs(x - np.floor(x)) - 0.5) < 1e-12))
  assert n_half >= 20           # the vectors do contain exact half-way cases (round-half-even matters)
  assert np.array_equal(ref_quant.duq_forward(w, -1.0, -1.0, 4), OPS["duq_passthrough_a"])
  assert np.array_equal(ref_quant.duq_forward(w, 1.0, 1.0, -1), OPS["duq_passthrough_bits"])


def test_duq_levels_c_oracle_vs_reference(oracle_lib):
  w = OPS["duq_w"]
  for case in META["duq_cases"]:
    L = 2 ** (case["bits"] - 1) - 1
    q = ref_int.duq_levels_c(w, None, F32(case["a"]), case["bits"]).astype(F32)
    assert np.array_equal(((q / F32(L)) * F32(case["c"])).astype(F32), OPS[case["key"]]), case


def test_prune_and_calibrators_vs_reference():
  assert np.array_equal(ref_quant.prune_forward(OPS["duq_w"], OPS["prune_mask"]), OPS["prune_out"])
  for k in ("randn", "zeros", "nonpos", "big"):
    x = OPS[f"cal_in_{k}"]
    for bits in (2, 8):
      for impl in (ref_quant, hq):        # oracle AND the product's host code
        assert F32(impl.gaussian_init(x, bits, True)) == OPS[f"cal_gauss_{k}_b{bits}"], (k, bits, impl.__name__)
        assert F32(impl.max_init(x, bits, True)) == OPS[f"cal_max_{k}_b{bits}"], (k, bits, impl.__name__)


def test_masks_vs_reference_source_lines():
  """Local and global magnitude masks + DuQ calibration produced by exec'ing
  examples/train_inpt_spikingjelly.py:147-229 on a toy tree (incl. a name that sorts out of numeric order)."""
  names = META["toy_names"]
  kernels = {n: OPS[f"toy_kernel_{n}"] for n in names}
  for impl in (ref_quant, hq):
    for p in (0.3, 0.75):
      gm = impl.global_masks(kernels, p)
      for n in names:
        assert np.array_equal(gm[n], OPS[f"toy_global_p{int(p * 100)}_{n}_mask"]), (impl.__name__, "global", p, n)
        assert np.array_equal(impl.local_mask(kernels[n], p), OPS[f"toy_local_p{int(p * 100)}_{n}_mask"]), (impl.__name__, p, n)
        assert F32(impl.gaussian_init(kernels[n], 8, True)) == OPS[f"toy_global_p{int(p * 100)}_{n}_a"][0]


def test_atan_and_lif_bit_exact_vs_reference(oracle_lib):
  assert np.array_equal(ref_snn.atan_fwd(OPS["atan_in"]), OPS["atan_out"])
  for ci, case in enumerate(META["lif_cases"]):
    x, u = OPS[f"lif{ci}_x"], OPS[f"lif{ci}_u0"].copy()
    for t in range(x.shape[0]):
      u, s = ref_snn.lif_step(u, x[t], case["tau"], case["v_threshold"], case["v_reset"])
      assert np.array_equal(u.view(np.uint32), OPS[f"lif{ci}_u"][t].view(np.uint32)), (ci, t)
      assert np.array_equal(s != 0, OPS[f"lif{ci}_s"][t] != 0), (ci, t)
  # case 1 holds exact ties un == 1: they spike (>= is inclusive) and reset to 0
  assert OPS["lif1_s"][0, :4].tolist() == [1, 1, 1, 1] and np.all(OPS["lif1_u"][0, :4] == 0)
  # the C oracle's LIF (folded fmaf form, zero carry) on the same pre-activations: scale 1, bias 0
  x = OPS["lif0_x"]
  s, uT = ref_int.lif_from_acc(x[:, :, None].copy(), np.ones(1, F32), np.zeros(1, F32))
  assert np.array_equal(s[:, :, 0], OPS["lif0_s"]) and np.array_equal(uT[:, 0], OPS["lif0_u"][-1])


def _layer_params(tag):
  srng = StableRNG(int(OPS[f"{tag}_seed"]))
  kshape = tuple(int(v) for v in OPS[f"{tag}_kshape"])
  k = (srng.standard_normal(kshape) * 0.3).astype(F32)
  m = (srng.uniform(0, 1, kshape) > 0.5).astype(F32)
  assert reffix.sha(k, m) == str(OPS[f"{tag}_ksha"])
  return {"kernel": k, "DuQ_0": {"a": OPS[f"{tag}_a"], "c": OPS[f"{tag}_c"]}, "prune_0": {"mask": m}}


def test_quantconv_quantdense_vs_reference(oracle_lib):
  """QuantConv 3x3 pad 1, QuantConv 1-D k=4 'SAME' (pads (1,2)) and QuantDense, 4 bit, a != c, masked."""
  lay = _layer_params("qconv3x3")
  y = ref_snn.quant_conv(lay, OPS["qconv3x3_x"].astype(F32), 4, ((1, 1), (1, 1)))
  assert np.max(np.abs(y - OPS["qconv3x3_y"])) <= 2e-6 * np.max(np.abs(OPS["qconv3x3_y"]))
  # integer path: exact accumulators x (c / L) reproduce the reference's float conv
  q = ref_int.duq_levels_c(lay["kernel"], lay["prune_0"]["mask"], lay["DuQ_0"]["a"][0], 4)
  acc = ref_int.conv3x3_acc(OPS["qconv3x3_x"], q)
  yi = acc.astype(np.float64) * (np.float64(lay["DuQ_0"]["c"][0]) / 7)
  assert np.max(np.abs(yi - OPS["qconv3x3_y"])) <= 2e-6 * np.max(np.abs(OPS["qconv3x3_y"]))
  lay = _layer_params("qconv1d")
  y = ref_snn.quant_conv(lay, OPS["qconv1d_x"], 4, "SAME")
  assert np.max(np.abs(y - OPS["qconv1d_y"])) <= 1e-6
  lay = _layer_params("qdense")
  y = ref_snn.quant_dense(lay, OPS["qdense_x"].astype(F32), 4)
  assert np.max(np.abs(y - OPS["qdense_y"])) <= 2e-6 * np.max(np.abs(OPS["qdense_y"]))
  q = ref_int.duq_levels_c(lay["kernel"], lay["prune_0"]["mask"], lay["DuQ_0"]["a"][0], 4)
  yi = ref_int.dense_acc(OPS["qdense_x"], q).astype(np.float64) * (np.float64(lay["DuQ_0"]["c"][0]) / 7)
  assert np.max(np.abs(yi - OPS["qdense_y"])) <= 2e-6 * np.max(np.abs(OPS["qdense_y"]))


def _oracle_vs_fixture(tag, float_path=True, forced=False):
  fx, m, v, fr = reffix.load_network(tag)
  ci = {}
  force = {}
  if forced:      # per-layer teacher forcing: every block sees the REFERENCE's spikes / attention as its input
    force = {"s1": reffix.ref_spikes(fx, "conv1"), "s2": reffix.ref_spikes(fx, "conv2"), "s3": reffix.ref_spikes(fx, "conv3"),
             "s4": reffix.ref_spikes(fx, "conv4"), "att4": fx["att4"], "s5": reffix.ref_spikes(fx, "conv5"),
             "att5": fx["att5"], "d1": reffix.ref_spikes(fx, "dense1")}
  li = ref_net.forward(ref_net.pack_network(v, m["bits"], m["H"]), fr, collect=ci, forced=force)
  keys = dict(conv1="s1", conv2="s2", conv3="s3", conv4="s4", conv5="s5", dense1="d1", dense2="d2")
  total = 0
  for n in reffix.BLOCKS:
    counts = None
    if n in ("conv4", "conv5"):
      counts = ci[keys[n]].sum(axis=(2, 3), dtype=np.int32)
    total += reffix.compare_block(fx, n, ci[keys[n]], u_final=ci[f"{n}_u"], counts_tbc=counts,
                                  upstream_flips=0 if forced else total)
  for k in ("att4", "att5"):
    assert np.max(np.abs(ci[k] - fx[k]) / fx[k]) <= 2e-6, k
  assert np.max(np.abs(li - fx["logits"])) <= reffix.logits_tolerance(m["T"], 10, total)
  if float_path:
    cf = {}
    lf = ref_snn.cextnet_forward(v, fr, m["bits"], collect=cf)
    kf = dict(conv1="pool1", conv2="pool2", conv3="pool3", conv4="conv4_spikes", conv5="conv5_spikes",
              dense1="dense1_spikes", dense2="dense2_spikes")
    tf = 0
    for n in reffix.BLOCKS:
      tf += reffix.compare_block(fx, n, (cf[kf[n]] != 0).astype(np.uint8), u_final=cf[f"{n}_u"], upstream_flips=tf)
    assert np.max(np.abs(lf - fx["logits"])) <= reffix.logits_tolerance(m["T"], 10, tf)
  return total


@pytest.mark.parametrize("tag", ["T4_H32_b8_p50", "T3_H32_b4_p80", "T3_H32_b2_p90", "T10_H32_b8_p50_c10"])
def test_oracle_network_vs_reference_cextnet(oracle_lib, tag):
  """Both oracle restatements against the reference's own CextNet.__call__ (examples/tcja/models.py:101-257):
  spikes of all seven blocks, final membranes, both attentions, logits."""
  _oracle_vs_fixture(tag)


@pytest.mark.parametrize("tag", ["T20_H128_b8_p50", "T20_H128_b4_p80", "T20_H128_b2_p90"])
def test_oracle_network_vs_reference_cextnet_full_size(oracle_lib, tag):
  """BASELINE.json configs[0] / [1] / [2] at their stated geometry (H = 128, T = 20): every block of the integer
  path, fed the reference's own spikes (teacher forcing per layer: a deep SNN amplifies a single boundary flip
  chaotically when free-running, which would hide what the flip budget is about), then the same free-running."""
  assert _oracle_vs_fixture(tag, float_path=False, forced=True) <= 8
  _oracle_vs_fixture(tag, float_path=False, forced=False)
