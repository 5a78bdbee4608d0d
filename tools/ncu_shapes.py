"""Replay selected GEMM launches of the training step in isolation for `ncu --set full --import-source on`.

Records the cb_gemm descriptors of one step (tools/autotune_gemm.record_step), picks the shapes whose key contains one of
--match, and launches each once on fresh operands between cudaProfilerStart/Stop (run under
`ncu --profile-from-start off`). Launch order = order of --match, so the i-th profiled kernel is the i-th pattern.
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch  # noqa: E402

import autotune_gemm as AT  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--match", nargs="+", required=True, help="substrings of ops.gemm_key (e.g. 'm401408 n256 k64 mode0 t1 r1')")
    for k, v in dict(batch=32, n_clips=2, n_frm=2, size=224, txt_len=32, n_ex=1, clip_batching=1).items():
        ap.add_argument("--" + k, type=int, default=v)
    args = ap.parse_args()
    from clipbert_b200 import ops
    ops.set_pdl(0)
    dev = torch.device("cuda:0")
    rec = AT.record_step(args)
    uniq = {}
    rec = [m for kw in rec for m in (kw["group"] if "group" in kw else [kw])]
    for kw in rec:
        uniq.setdefault(ops.gemm_key(kw), kw)
    del rec
    torch.cuda.empty_cache()
    ops._tuning = None            # replay with the production launch configuration (tuning table included)
    os.environ.pop("CB_NO_TUNING", None)
    calls = []
    for pat in args.match:
        hits = [k for k in uniq if pat in k]
        if not hits:
            print("no recorded GEMM matches %r" % pat)
            continue
        calls.append((hits[0], AT.synth_call(uniq[hits[0]], dev)))
    for key, call in calls:      # warm-up (tensor maps, L2 state comparable to a mid-step launch)
        ops.gemm(**call)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for key, call in calls:
        ops.gemm(**call)
        torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    for i, (key, _) in enumerate(calls):
        print("profiled launch %d: %s" % (i, key))


if __name__ == "__main__":
    main()
