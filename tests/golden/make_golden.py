"""Generate the committed golden fixtures from the CPU oracle.

    python tests/golden/make_golden.py

Round-1 fixtures, kept as a second line of defence: these vectors come from the
oracle restatements -- the float path in the reference's op order
(oracle/ref_snn.py) and the integer path (oracle/ref_net.py), which must agree
before anything is written.  The fixtures that pin the oracle itself to the
executed reference are written by tests/golden/make_from_reference.py."""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_events, ref_int, ref_net, ref_quant, ref_snn  # noqa: E402
from snnquantprune_b200 import synthetic  # noqa: E402


def sha(*arrays) -> str:
  h = hashlib.sha256()
  for a in arrays:
    h.update(np.ascontiguousarray(a).tobytes())
  return h.hexdigest()


def variables_digest(v) -> str:
  arrs = []
  for name in sorted(v["params"].keys()):
    lay = v["params"][name]
    if "kernel" in lay:
      arrs += [lay["kernel"], lay["prune_0"]["mask"], lay["DuQ_0"]["a"], lay["DuQ_0"]["c"]]
    else:
      arrs += [lay["scale"], lay["bias"], v["batch_stats"][name]["mean"], v["batch_stats"][name]["var"]]
  return sha(*arrs)


def duq_vectors():
  rng = np.random.default_rng(11)
  w = np.concatenate([rng.standard_normal(61).astype(np.float32) * 0.3,
                      np.array([0.0, 1e-9, -1e-9, 0.5, -0.5, 2.0, -2.0], np.float32)])
  # exact half-way cases for round-half-even at L = 7 and L = 127 with a = 1
  w = np.concatenate([w, np.array([0.5 / 7, 1.5 / 7, 2.5 / 7, -0.5 / 7, 0.5 / 127, 1.5 / 127], np.float32)])
  mask = (rng.uniform(size=w.shape) > 0.3).astype(np.float32)
  out = {"w": w.tolist(), "mask": mask.tolist(), "cases": []}
  for bits in (2, 4, 8):
    for a, c in ((1.0, 1.0), (0.37, 0.41)):
      q = ref_quant.duq_levels(w, a, bits)
      wq = ref_quant.effective_weight(w, a, c, mask, bits)
      out["cases"].append({"bits": bits, "a": a, "c": c, "levels": q.tolist(),
                           "forward": [float(x) for x in wq]})
  return out


def network_fixture(bits, p, T, H, B, seed_w, seed_x):
  v = synthetic.make_variables(bits=bits, prune_percentage=p, T=T, H=H, seed=seed_w)
  fr = synthetic.make_frames(B, T, H, H, seed=seed_x)
  pk = ref_net.pack_network(v, bits, H)
  ci, cf = {}, {}
  logits_int = ref_net.forward(pk, fr, collect=ci)
  logits_f = ref_snn.cextnet_forward(v, fr, bits, collect=cf)
  # the two restatements must agree before a fixture is written
  pairs = (("s1", "pool1"), ("s2", "pool2"), ("s3", "pool3"), ("s4", "conv4_spikes"),
           ("s5", "conv5_spikes"), ("d1", "dense1_spikes"), ("d2", "dense2_spikes"))
  flips = {a: float(np.mean(ci[a] != (cf[b] != 0))) for a, b in pairs}
  assert max(flips.values()) <= 1e-4, flips
  assert np.abs(logits_int - logits_f).max() <= 2e-2
  fx = {
      "meta": np.array(json.dumps(dict(bits=bits, prune=p, T=T, H=H, B=B, seed_w=seed_w,
                                       seed_x=seed_x, variables_sha=variables_digest(v),
                                       frames_sha=sha(fr), flips_int_vs_float=flips,
                                       rates=ref_net.firing_rates(ci)))),
      "logits_int": logits_int, "logits_float": logits_f,
      "att4": ci["att4"], "att5": ci["att5"],
      "conv5_u": ci["conv5_u"], "dense2_u": ci["dense2_u"],
      "conv2_acc_sum": np.array([ci["conv2_acc"].astype(np.int64).sum(),
                                 (ci["conv2_acc"].astype(np.int64) ** 2).sum()]),
  }
  for k in ("s1", "s2", "s3", "s4", "s5", "d1", "d2"):
    fx[k + "_bits"] = np.packbits(ci[k].reshape(-1))
    fx[k + "_shape"] = np.array(ci[k].shape)
  return fx


def events_fixture():
  """Ragged batch of event lists (one empty, one shorter than T, one with a hot pixel that saturates uint8 and
  events beyond the row / frame) -> frames of the oracle's preprocess_data_number, for two resolution scales,
  plus the sowed densities of the uint8 frames."""
  rng = np.random.default_rng(21)
  wh, T = 32, 6
  sizes = [4000, 0, 4, 2500]
  samples = []
  for n in sizes:
    a = np.stack([rng.integers(0, wh, n), rng.integers(0, wh, n), rng.integers(0, 3, n)], 1).astype(np.int32)
    samples.append(a)
  samples[0][:600, :2] = [7, 3]                       # hot pixel: > 255 events in one frame
  samples[3][::50, 0] = wh + 1                        # beyond the row: lands in the next row (flat position)
  samples[3][-3:, 1] = wh + 2                         # beyond the frame: dropped
  fx = {"addrs": np.concatenate(samples, 0), "offsets": np.cumsum([0] + sizes).astype(np.int64),
        "wh": np.array(wh), "T": np.array(T)}
  for rs in (1, 2):
    f32 = ref_events.batch_to_frames(samples, T, wh, rs)
    f8, nsat = ref_events.batch_to_frames(samples, T, wh, rs, saturate_u8=True)
    fx[f"frames_i32_rs{rs}"] = f32
    fx[f"n_saturated_rs{rs}"] = np.array(nsat)
    d = ref_events.sow_densities(np.swapaxes(f8, 0, 1))
    fx[f"density_counts_rs{rs}"] = d["counts"]
  return fx


if __name__ == "__main__":
  with open(os.path.join(HERE, "duq_vectors.json"), "w") as f:
    json.dump(duq_vectors(), f)
  np.savez_compressed(os.path.join(HERE, "events_T6_wh32.npz"), **events_fixture())
  np.savez_compressed(os.path.join(HERE, "cextnet_T4_H32_b8_p50.npz"),
                      **network_fixture(8, 0.5, 4, 32, 2, 1, 0))
  np.savez_compressed(os.path.join(HERE, "cextnet_T3_H32_b4_p80.npz"),
                      **network_fixture(4, 0.8, 3, 32, 2, 5, 7))
  np.savez_compressed(os.path.join(HERE, "cextnet_T3_H32_b2_p90.npz"),
                      **network_fixture(2, 0.9, 3, 32, 1, 6, 8))
  print("golden fixtures written")
