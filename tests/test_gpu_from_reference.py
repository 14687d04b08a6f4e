"""GPU parity against fixtures made by EXECUTING THE UNMODIFIED REFERENCE
(tests/golden/make_from_reference.py).  Everything goes through the C-ABI.

* quantizer / prune / pack kernels: bit-exact;
* LIF: a fused block on a grid of exactly representable pre-activations, bit-exact
  (generic and production epilogues);
* whole network, small geometries: free-running, every block;
* BASELINE.json configs[0]/[1]/[2] at H = 128, T = 20: every block fed the
  reference's own spikes (per-layer teacher forcing), then free-running logits."""
import json
import os

import numpy as np
import pytest
import torch

import reffix
from oracle import ref_int
from snnquantprune_b200 import _lib
from snnquantprune_b200._lib import BlockParams

pytestmark = pytest.mark.gpu
F32 = np.float32
DEV = "cuda"
OPS = np.load(os.path.join(reffix.GOLD, "from_reference_ops.npz"))
META = json.loads(str(OPS["meta"]))


def dev(a, dtype=None):
  t = torch.as_tensor(np.ascontiguousarray(a), device=DEV)
  return t if dtype is None else t.to(dtype)


P = _lib.ptr


def bt(a):
  """fixture (T,B,...) -> engine layout (B,T,...), contiguous."""
  return np.ascontiguousarray(np.swapaxes(a, 0, 1))


def test_duq_prune_pack_kernels_vs_reference(cuda_lib):
  w = OPS["duq_w"]
  wd = dev(w)
  st = _lib.stream()
  for case in META["duq_cases"]:
    ref = OPS[case["key"]]
    ad, cd = dev(np.array([case["a"]], F32)), dev(np.array([case["c"]], F32))
    out = torch.empty_like(wd)
    _lib.check(cuda_lib.snnqp_duq_forward(P(wd), None, P(ad), P(cd), case["bits"], w.size, P(out), st))
    assert np.array_equal(out.cpu().numpy().view(np.uint32), ref.view(np.uint32)), case
    q = torch.empty(w.shape, device=DEV, dtype=torch.int8)
    _lib.check(cuda_lib.snnqp_pack_levels(P(wd), None, P(ad), case["bits"], w.size, P(q), st))
    L = 2 ** (case["bits"] - 1) - 1
    assert np.array_equal(((q.cpu().numpy().astype(F32) / F32(L)) * F32(case["c"])).astype(F32), ref), case
  # prune: DuQ pass-through (bits = -1) leaves the mask multiply (quant.py:475-491)
  md = dev(OPS["prune_mask"])
  one = dev(np.array([1.0], F32))
  out = torch.empty_like(wd)
  _lib.check(cuda_lib.snnqp_duq_forward(P(wd), P(md), P(one), P(one), -1, w.size, P(out), st))
  assert np.array_equal(out.cpu().numpy(), OPS["prune_out"])


@pytest.mark.parametrize("impl", [_lib.IMPL_TCGEN05, _lib.IMPL_SIMT])
def test_lif_bit_exact_vs_reference_multi_step_lif(cuda_lib, impl):
  """multi_step_LIF (spiking_learning.py:404-416) + atan (:221-224) inside the fused 3x3 block: identity centre
  tap, scale 2^-6, bias -0.25 => the block's pre-activation is exactly the fixture's x, the carry starts at 0
  (initialize_carry, :464-472).  Generic epilogue: un-pooled spikes and the final membrane, bit for bit;
  production epilogue: pooled spikes."""
  cnt = OPS["lifgrid_counts"]                          # (T,1,H,W,C)
  T, B, H, W, C = cnt.shape
  ref_s = np.unpackbits(OPS["lifgrid_bits"])[:cnt.size].reshape(cnt.shape)
  wq = np.zeros((9, C, C), np.int8)
  wq[4] = np.eye(C, dtype=np.int8)                     # centre tap, [tap][cout][cin]
  blob = torch.zeros((int(cuda_lib.snnqp_conv3x3_blob_bytes(C, C)),), device=DEV, dtype=torch.int8)
  blob[:9 * C * C] = dev(wq).reshape(-1)
  nz = torch.empty((36,), device=DEV, dtype=torch.uint8)
  _lib.check(cuda_lib.snnqp_conv3x3_slab_bitmap(P(blob), C, C, P(nz), _lib.stream()))
  blob[9 * C * C:9 * C * C + 36] = nz.view(torch.int8)
  scale, bias = dev(np.full(C, 2.0 ** -6, F32)), dev(np.full(C, -0.25, F32))
  x = dev(bt(cnt))
  for pool in (0, 1):
    p = BlockParams()
    p.T, p.B, p.H, p.W, p.Cin, p.Cout = T, B, H, W, C, C
    p.x_stride_b, p.x_stride_t = x.stride(0), x.stride(1)
    y = torch.empty((B, T, H >> pool, W >> pool, C), device=DEV, dtype=torch.uint8)
    p.y_stride_b, p.y_stride_t = y.stride(0), y.stride(1)
    p.tau, p.v_threshold, p.v_reset, p.pool, p.impl = 2.0, 1.0, 0.0, pool, impl
    u = torch.empty((B, H, W, C), device=DEV, dtype=torch.float32) if pool == 0 else None
    _lib.check(cuda_lib.snnqp_spiking_conv3x3_fwd(p, P(x), None, P(blob), P(scale), P(bias), P(y), P(u), None, _lib.stream()))
    got = np.swapaxes(y.cpu().numpy(), 0, 1)
    if pool == 0:
      assert np.array_equal(got, ref_s)
      assert np.array_equal(u.cpu().numpy().view(np.uint32), OPS["lifgrid_uT"].view(np.uint32))
    else:
      assert np.array_equal(got, ref_int.maxpool2_u8(ref_s))


def _engine(v, m, chunk=16, impl=_lib.IMPL_AUTO, lif_mode=_lib.LIF_EXACT):
  from snnquantprune_b200 import CextNetEngine, pack_cextnet
  return CextNetEngine(pack_cextnet(v, m["bits"], m["T"], m["H"], num_classes=m["num_classes"], device=DEV),
                       impl=impl, chunk=chunk, lif_mode=lif_mode)


@pytest.mark.parametrize("tag", ["T4_H32_b8_p50", "T3_H32_b4_p80", "T3_H32_b2_p90", "T10_H32_b8_p50_c10"])
def test_network_free_running_vs_reference_cextnet(cuda_lib, tag):
  """CUDA forward (instrumented: every intermediate) against the reference's own CextNet.__call__."""
  fx, m, v, fr = reffix.load_network(tag)
  eng = _engine(v, m)
  c = {}
  logits = eng.forward(dev(fr), collect=c).cpu().numpy()
  tb = lambda k: np.swapaxes(c[k].cpu().numpy(), 0, 1)
  keys = dict(conv1="s1", conv2="s2", conv3="s3", conv4="s4", conv5="s5", dense1="d1", dense2="d2")
  total = 0
  for n in reffix.BLOCKS:
    total += reffix.compare_block(fx, n, tb(keys[n]), u_final=c[f"{n}_u"].cpu().numpy(), upstream_flips=total)
  for k in ("att4", "att5"):
    assert np.max(np.abs(tb(k) - fx[k]) / fx[k]) <= 2e-6, k
  assert np.max(np.abs(logits - fx["logits"])) <= reffix.logits_tolerance(m["T"], 10, total)
  # production path (fused tail, no instrumentation): same logits as the instrumented pass
  assert np.array_equal(eng.forward(dev(fr)).cpu().numpy(), logits)


@pytest.mark.parametrize("tag", ["T20_H128_b8_p50", "T20_H128_b4_p80", "T20_H128_b2_p90"])
def test_full_size_layerwise_vs_reference_cextnet(cuda_lib, tag):
  """BASELINE.json configs[0] / [1] / [2] at H = 128, T = 20.  Every block runs on the REFERENCE's spikes and
  attention (teacher forcing per layer) through the same engine launches the production forward uses; flips are
  counted per block against the 1e-4 budget, final membranes to 1e-5, attention to 2e-6.  Then free-running."""
  fx, m, v, fr = reffix.load_network(tag)
  eng = _engine(v, m)
  pk = eng.pk
  T, H, C, B = pk.T, pk.H, pk.channels, m["B"]
  u8 = dict(device=DEV, dtype=torch.uint8)
  rs = {n: reffix.ref_spikes(fx, n) for n in reffix.BLOCKS}
  flips = {}

  def conv_block(i, name, x, Hin, Cin, pool, att=None):
    c = {}
    y = torch.empty((B, T, Hin >> pool, Hin >> pool, C), **u8)
    eng._conv(i, x, y, B, Hin, Cin, pool, att=att, collect=c, key=name)         # instrumented launch: + membranes
    flips[name] = reffix.compare_block(fx, name, np.swapaxes(y.cpu().numpy(), 0, 1), u_final=c[name + "_u"].cpu().numpy())
    return y

  def conv_block_fast(i, name, x, Hin, Cin, att=None, counts=None):
    y = torch.empty((B, T, Hin // 2, Hin // 2, C), **u8)
    eng._conv(i, x, y, B, Hin, Cin, 1, att=att, counts=counts)                  # production launch
    return y

  # conv1-3: production launch (pooled spikes) + instrumented launch (un-pooled is not stored; membranes are)
  x = dev(fr)
  for i, (name, Hin, Cin) in enumerate((("conv1", H, 2), ("conv2", H // 2, C), ("conv3", H // 4, C))):
    y = conv_block_fast(i, name, x, Hin, Cin)
    flips[name] = reffix.compare_block(fx, name, np.swapaxes(y.cpu().numpy(), 0, 1))
    # bit-packed spikes in and out (SNNQP_SPIKES_BITS: the production layout between conv1 .. conv4): same result
    xb = x if Cin == 2 else dev(np.packbits(x.cpu().numpy(), axis=-1, bitorder="little"))
    yb = torch.empty((B, T, Hin // 2, Hin // 2, C // 8), **u8)
    eng._conv(i, xb, yb, B, Hin, Cin, 1)
    assert np.array_equal(np.unpackbits(yb.cpu().numpy(), axis=-1, bitorder="little"), y.cpu().numpy()), name + " (bits)"
    c = {}
    yu = torch.empty((B, T, Hin, Hin, C), **u8)
    eng._conv(i, x, yu, B, Hin, Cin, 0, collect=c, key=name)
    assert np.array_equal(ref_int.maxpool2_u8(yu.cpu().numpy()), y.cpu().numpy()), name   # both epilogues agree
    reffix.compare_block(fx, name, np.swapaxes(y.cpu().numpy(), 0, 1), u_final=c[name + "_u"].cpu().numpy())
    if i == 0:
      # conv1 with the leak on the tensor core (SNNQP_LIF_TENSOR, the engine's default): tolerance parity against the
      # executed reference -- pooled spikes under the 1e-4 flip budget, final membranes to 1e-5
      eng_t = _engine(v, m, lif_mode=_lib.LIF_TENSOR)
      for packed_out in (False, True):
        yt = torch.empty((B, T, Hin // 2, Hin // 2, C // 8 if packed_out else C), **u8)
        ct = {}
        eng_t._conv(0, x, yt, B, Hin, Cin, 1, collect=ct, key=name, collect_acc=False)
        st = np.unpackbits(yt.cpu().numpy(), axis=-1, bitorder="little") if packed_out else yt.cpu().numpy()
        flips["conv1_lif_tensor"] = reffix.compare_block(fx, name, np.swapaxes(st, 0, 1), u_final=ct[name + "_u"].cpu().numpy())
    x = dev(bt(rs[name]))                                                       # next block sees the reference's spikes
  # conv4 (+ spike counts) -> TCJA
  s4 = conv_block(3, "conv4", x, H // 8, C, 0)
  cnt = torch.zeros((B, T, C), device=DEV, dtype=torch.int32)
  p4 = conv_block_fast(3, "conv4", x, H // 8, C, counts=cnt)
  assert np.array_equal(p4.cpu().numpy(), ref_int.maxpool2_u8(s4.cpu().numpy()))
  assert np.array_equal(cnt.cpu().numpy(), s4.cpu().numpy().sum(axis=(2, 3), dtype=np.int32))
  cnt2 = torch.zeros_like(cnt)                 # bit-packed input, u8 output + counts: what the production tail runs
  p4b = conv_block_fast(3, "conv4", dev(np.packbits(x.cpu().numpy(), axis=-1, bitorder="little")), H // 8, C, counts=cnt2)
  assert np.array_equal(p4b.cpu().numpy(), p4.cpu().numpy()) and np.array_equal(cnt2.cpu().numpy(), cnt.cpu().numpy())
  s4r = dev(bt(rs["conv4"]))
  cnt_r = dev(bt(fx["conv4_counts"]))
  att = torch.empty((B, T, C), device=DEV, dtype=torch.float32)
  eng._tcja(0, B, H // 8, None, cnt_r, att)
  assert np.max(np.abs(np.swapaxes(att.cpu().numpy(), 0, 1) - fx["att4"]) / fx["att4"]) <= 2e-6
  # conv5: real-valued input att4 * pool(s4), both from the reference
  p4r = dev(ref_int.maxpool2_u8(bt(rs["conv4"])))
  att4r = dev(bt(fx["att4"]))
  conv_block(4, "conv5", p4r, H // 16, C, 0, att=att4r)
  cnt_r = dev(bt(fx["conv5_counts"]))
  eng._tcja(1, B, H // 16, None, cnt_r, att)
  assert np.max(np.abs(np.swapaxes(att.cpu().numpy(), 0, 1) - fx["att5"]) / fx["att5"]) <= 2e-6
  # dense1 (att5 * pool(s5), flatten folded into the weights), dense2, vote
  p5r = dev(ref_int.maxpool2_u8(bt(rs["conv5"])))
  att5r = dev(bt(fx["att5"]))
  for lay, name, xin, a in ((pk.dense1, "dense1", p5r.view(B, T, -1), att5r), (pk.dense2, "dense2", dev(bt(rs["dense1"])), None)):
    c = {}
    y = torch.empty((B, T, lay.cout), **u8)
    eng._dense(lay, B, xin, a, y, c, name)
    flips[name] = reffix.compare_block(fx, name, np.swapaxes(y.cpu().numpy(), 0, 1), u_final=c[name + "_u"].cpu().numpy())
  assert sum(flips.values()) <= 8, flips
  # free-running production forward (bit-packed spikes; reference-order LIF, then the single-rounding and the
  # tensor-core-leak conv1 -- the engine's default): logits within the flip-derived tolerance of the reference's
  # (LIF_TENSOR flips ~1e-8 .. 1.5e-7 of conv1's spikes -- up to one or two per sample -- and a free-running SNN with
  # random weights amplifies each: the logits move by a few quanta of 1 / (T * 10) = 0.005; tools/probe_lif_modes_logits.py)
  top2 = np.sort(fx["logits"], -1)[:, -2:]
  for lm, tol in ((_lib.LIF_EXACT, 0.02), (_lib.LIF_FAST, 0.02), (_lib.LIF_TENSOR, 0.08)):
    logits = _engine(v, m, chunk=296, lif_mode=lm).forward(dev(fr)).cpu().numpy()
    assert np.max(np.abs(logits - fx["logits"])) <= tol, (lm, logits, fx["logits"])
    clear = (top2[:, 1] - top2[:, 0]) > 2 * tol           # argmax is only defined up to the tolerance
    assert np.array_equal(np.argmax(logits, -1)[clear], np.argmax(fx["logits"], -1)[clear])


def test_production_shape_chunk_296(cuda_lib, oracle_lib):
  """The benchmarked configuration: batch >= chunk = 296 (a full wave-quantised head chunk plus a ragged one).
  Logits must be identical to the chunk = 16 schedule, and the first / last samples must match the oracle."""
  from oracle import ref_net
  from snnquantprune_b200 import synthetic
  bits, T, H, B = 8, 20, 128, 300
  v = synthetic.make_variables(bits=bits, prune_percentage=0.5, T=T, H=H, seed=1, stable=True)
  fr = synthetic.make_frames(B, T, H, H, seed=77, stable=True)
  m = dict(bits=bits, T=T, H=H, num_classes=11)
  frd = dev(fr)
  l296 = _engine(v, m, chunk=296).forward(frd).cpu().numpy()
  l16 = _engine(v, m, chunk=16).forward(frd).cpu().numpy()
  assert np.array_equal(l296, l16)
  # the bench's configuration: single-rounding LIF in conv1.  ~4e-9 of its 3.1e9 pooled spikes flip (a handful in this
  # batch); a flipped spike perturbs that sample's deeper layers, so a few samples move by a few output spikes
  lfast = _engine(v, m, chunk=296, lif_mode=_lib.LIF_FAST).forward(frd).cpu().numpy()
  changed = np.any(np.abs(lfast - l296) > 1e-6, axis=1)
  assert changed.mean() <= 0.10 and np.max(np.abs(lfast - l296)) <= 0.05       # <= 10 output spikes of T * 10 = 200
  assert np.mean(lfast.argmax(-1) == l296.argmax(-1)) >= 0.98
  # the engine's default: conv1's leak on the tensor core (~1.5e-8 of the pooled spikes flip)
  # on this weight set 1.5e-7 of conv1's pooled spikes, ~1.5 per sample: most samples see at least one flip, each
  # amplified by the free-running layers behind it into a few output spikes (quantum 1 / 200)
  eng_t = _engine(v, m, chunk=296, lif_mode=_lib.LIF_TENSOR)
  ltens = eng_t.forward(frd).cpu().numpy()
  dl = np.abs(ltens - l296).max(axis=1)
  assert dl.max() <= 0.10 and dl.mean() <= 0.03, (float(dl.max()), float(dl.mean()))      # measured 0.06 / 0.022
  top2 = np.sort(l296, -1)[:, -2:]
  clear = (top2[:, 1] - top2[:, 0]) > 0.05
  assert np.mean(ltens.argmax(-1)[clear] == l296.argmax(-1)[clear]) >= 0.97 and np.mean(ltens.argmax(-1) == l296.argmax(-1)) >= 0.88
  n = 296                                              # conv1 alone: the flip budget itself
  s_t = eng_t._workspace(n, n)["s1"][:n]
  eng_e = _engine(v, m, chunk=296)
  s_e = eng_e._workspace(n, n)["s1"][:n]
  eng_t._conv(0, frd[:n], s_t, n, H, 2, 1)
  eng_e._conv(0, frd[:n], s_e, n, H, 2, 1)
  flips = int((eng_t.unpack_spikes(s_t) != eng_e.unpack_spikes(s_e)).sum().item())
  assert flips <= 1e-4 * s_t.numel() * 8 and flips <= 1e-6 * s_t.numel() * 8, flips      # bar 1e-4; measured 1.5e-7
  pkd = ref_net.pack_network(v, bits, H)
  idx = [0, 295, 299]
  lo = ref_net.forward(pkd, fr[idx])
  assert np.max(np.abs(l296[idx] - lo)) <= 0.02 and np.array_equal(np.argmax(l296[idx], -1), np.argmax(lo, -1))


def _layer_params(tag):
  from snnquantprune_b200.synthetic import StableRNG
  srng = StableRNG(int(OPS[f"{tag}_seed"]))
  kshape = tuple(int(v) for v in OPS[f"{tag}_kshape"])
  k = (srng.standard_normal(kshape) * 0.3).astype(F32)
  m = (srng.uniform(0, 1, kshape) > 0.5).astype(F32)
  assert reffix.sha(k, m) == str(OPS[f"{tag}_ksha"])
  return {"params": {"kernel": k, "DuQ_0": {"a": OPS[f"{tag}_a"], "c": OPS[f"{tag}_c"]}, "prune_0": {"mask": m}}}


def test_layer_facades_vs_reference_quantconv_quantdense(cuda_lib):
  """The drop-in layer facades called on their own (same constructor fields as the reference's modules) against the
  outputs of the reference's QuantConv.__call__ (3x3 pad 1 and 1-D k = 4 'SAME') and QuantDense.__call__."""
  from snnquantprune_b200 import QuantConv, QuantDense, QuantConfig
  cfg = QuantConfig(bits=4, g_scale=0.0, prune_percentage=0.5)
  y = QuantConv(features=128, kernel_size=(3, 3), padding=((1, 1), (1, 1)), use_bias=False, config=cfg, bits=4,
                g_scale=0.0).apply(_layer_params("qconv3x3"), dev(OPS["qconv3x3_x"]))
  ref = OPS["qconv3x3_y"]
  assert np.max(np.abs(y.cpu().numpy() - ref)) <= 2e-6 * np.max(np.abs(ref))
  y = QuantConv(features=5, kernel_size=[4], padding="SAME", use_bias=False, config=cfg, bits=4,
                g_scale=0.0).apply(_layer_params("qconv1d"), dev(OPS["qconv1d_x"]))
  assert y.shape == OPS["qconv1d_y"].shape and np.max(np.abs(y.cpu().numpy() - OPS["qconv1d_y"])) <= 1e-6
  y1 = QuantConv(features=5, kernel_size=4, padding="SAME", use_bias=False, config=cfg, bits=4,
                 g_scale=0.0).apply(_layer_params("qconv1d"), dev(OPS["qconv1d_x"][0]))       # single input (W, Cin)
  assert np.array_equal(y1.cpu().numpy(), y[0].cpu().numpy())
  for x in (dev(OPS["qdense_x"]), dev(OPS["qdense_x"].astype(F32))):
    y = QuantDense(110, use_bias=False, config=cfg, bits=4, g_scale=0.0).apply(_layer_params("qdense"), x)
    ref = OPS["qdense_y"]
    assert np.max(np.abs(y.cpu().numpy() - ref)) <= 2e-6 * np.max(np.abs(ref))
  with pytest.raises(NotImplementedError):
    QuantConv(features=5, kernel_size=[4], padding="VALID", use_bias=False, config=cfg).apply(
        _layer_params("qconv1d"), dev(OPS["qconv1d_x"]))


@pytest.mark.parametrize("bits,prune", [(4, 0.8), (2, 0.9)])
def test_full_size_configs_1_2_free_running_two_samples(cuda_lib, oracle_lib, bits, prune):
  """BASELINE.json configs[1] (4 bit / 80 %) and configs[2] (2 bit / 90 %) at H = 128, T = 20 with B = 2, nothing
  teacher-forced: the production forward (bit-packed spikes, pad-free tiles, fused tail) against the integer-path
  oracle -- itself pinned to the executed reference at this geometry (test_from_reference_cpu.py).  Integer-input
  blocks must agree bit for bit; dense1 (real-valued input, free-running) within the flip budget; logits within the
  quantum of the counted flips."""
  from oracle import ref_net
  from snnquantprune_b200 import synthetic
  T, H, B = 20, 128, 2
  v = synthetic.make_variables(bits=bits, prune_percentage=prune, T=T, H=H, seed=5 + bits, stable=True)
  fr = synthetic.make_frames(B, T, H, H, seed=7 + bits, stable=True)
  m = dict(bits=bits, T=T, H=H, num_classes=11)
  eng = _engine(v, m, chunk=296)
  c = {}
  li = eng.forward(dev(fr), collect=c).cpu().numpy()                    # instrumented pass (every intermediate)
  lp = eng.forward(dev(fr)).cpu().numpy()                               # production pass
  assert np.array_equal(li, lp)
  co = {}
  lo = ref_net.forward(ref_net.pack_network(v, bits, H), fr, collect=co)
  tb = lambda k: np.swapaxes(c[k].cpu().numpy(), 0, 1)
  for k in ("s1", "s2", "s3", "s4"):
    assert np.array_equal(tb(k), co[k]), k
  flips = 0
  for k in ("s5", "d1", "d2"):
    d = int((tb(k) != co[k]).sum())
    assert d <= 1e-4 * co[k].size + 8 * flips, (k, d)
    flips += d
  assert np.max(np.abs(lp - lo)) <= reffix.logits_tolerance(T, 10, flips)
  assert min(float(co[k].mean()) for k in ("s1", "s2", "s3", "s4", "s5", "d1", "d2")) > 0.005      # every block fires
