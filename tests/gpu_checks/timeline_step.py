"""In-graph kernel timeline of one LECO iteration (CUPTI through torch.profiler, no ncu): the launch lists under
profiles/ are ncu passes (serialised, cold cache), this one shows what each kernel costs INSIDE the replayed CUDA graphs
(warm L2, programmatic dependent launch overlapping prologues).  Per kernel name: launches, summed duration, and
"cost" = by how much the kernel extends the timeline past everything launched before it (end-to-end deltas: correct
under PDL overlap), for (a) the first CFG denoise step (up to guided_step_kernel) and (b) the whole iteration.

  python tests/gpu_checks/timeline_step.py --k 2 --out gpurun_out/timeline_sd21.md
"""
import argparse
import contextlib
import io
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def short(name: str) -> str:
    name = name.replace("leco::", "")
    cut = name.find("(")
    return (name if cut < 0 else name[:cut])[:70]


def table(events, title):
    """Attribution by END times: under programmatic dependent launch a kernel starts (and its CUPTI duration begins)
    while its predecessor is still running, so start-to-start or raw durations charge the predecessor's run time to
    the wrong kernel.  cost_i = end_i - max(end of everything before it, start_i) = by how much kernel i extends the
    timeline; time with no kernel resident is reported as idle (host gaps between graph replays)."""
    if not events:
        return f"### {title}\n(no kernels)\n"
    agg = {}
    prev_end, idle = events[0]["ts"], 0.0
    for e in events:
        st, en = e["ts"], e["ts"] + e["dur"]
        if st > prev_end:
            idle += st - prev_end
        cost = max(en - max(prev_end, st), 0.0)
        prev_end = max(prev_end, en)
        a = agg.setdefault(short(e["name"]), [0, 0.0, 0.0])
        a[0] += 1
        a[1] += e["dur"]
        a[2] += cost
    span = prev_end - events[0]["ts"]
    rows = [f"### {title}\n", f"launches: {len(events)}   span: {span / 1e3:.3f} ms   idle (no kernel resident): {idle / 1e3:.3f} ms"
            f"   busy: {(span - idle) / 1e3:.3f} ms\n",
            "| kernel | launches | cost ms | share of busy | avg cost us | avg CUPTI duration us |", "|---|---|---|---|---|---|"]
    for name, (n, dur, cost) in sorted(agg.items(), key=lambda kv: -kv[1][2]):
        rows.append(f"| `{name}` | {n} | {cost / 1e3:.3f} | {100 * cost / max(span - idle, 1e-9):.1f}% | {cost / n:.1f} | {dur / n:.1f} |")
    return "\n".join(rows) + "\n"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--k", type=int, default=2)
    ap.add_argument("--arch", default="sd21")
    ap.add_argument("--res", type=int, default=512)
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--rank", type=int, default=4)
    ap.add_argument("--c3lier", action="store_true", help="attention + conv adapters (train_lora.py:44-46)")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "timeline.md"))
    args = ap.parse_args()
    import torch
    from torch.profiler import ProfilerActivity, profile
    from leco_b200 import lora as plora
    from leco_b200.lora import LoRANetwork
    from leco_b200.scheduler import DDIMScheduler
    from leco_b200.synthetic import build_engine, prompt_embedding
    from leco_b200.trainer import LecoTrainer, PromptPair
    from leco_b200.unet import SPECS
    dev = torch.device("cuda", 0)
    unet = build_engine(args.arch, dev, seed=0)
    torch.manual_seed(1234)
    saved = list(plora.DEFAULT_TARGET_REPLACE)
    try:
        if args.c3lier:
            plora.DEFAULT_TARGET_REPLACE += plora.UNET_TARGET_REPLACE_MODULE_CONV
        with contextlib.redirect_stdout(io.StringIO()):
            net = LoRANetwork(unet, rank=args.rank, multiplier=1.0, alpha=1.0, train_method="full")
    finally:
        plora.DEFAULT_TARGET_REPLACE[:] = saved
    net.to(dev, dtype=torch.bfloat16)
    D = SPECS[args.arch].cross_attention_dim
    emb = {p: prompt_embedding(p, D) for p in ("van gogh", "")}
    pair = PromptPair(target=emb["van gogh"], positive=emb["van gogh"], unconditional=emb[""], neutral=emb[""],
                      guidance_scale=1.0, resolution=args.res, batch_size=args.batch, action="erase")
    v_pred = "v_prediction" if args.arch in ("sd21", "tiny21") else "epsilon"
    tr = LecoTrainer(unet, net, DDIMScheduler(v_pred), [pair], lr=1e-4, max_denoising_steps=50)
    for _ in range(3):
        tr.iteration(fixed_k=args.k)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        tr.iteration(fixed_k=args.k)
        torch.cuda.synchronize()
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "trace.json")
        prof.export_chrome_trace(path)
        trace = json.load(open(path))
    ev = sorted((e for e in trace["traceEvents"] if e.get("cat") == "kernel"), key=lambda e: e["ts"])
    names = sorted({e["name"] for e in ev})
    with open(os.path.splitext(args.out)[0] + "_events.json", "w") as f:   # raw list for offline analysis
        json.dump({"names": names, "events": [[names.index(e["name"]), round(e["ts"], 3), round(e["dur"], 3),
                                               e.get("args", {}).get("grid", None)] for e in ev]}, f)
    first = next((i for i, e in enumerate(ev) if "guided_step" in e["name"] or "sched_step" in e["name"]), len(ev) - 1)
    second = next((i for i in range(first + 1, len(ev)) if "guided_step" in ev[i]["name"] or "sched_step" in ev[i]["name"]),
                  first) if args.k > 1 else first
    md = [f"# In-graph kernel timeline: {args.arch}, batch {args.batch}, {args.res} px, k = {args.k} "
          f"(torch.profiler / CUPTI, one iteration after 3 warm-up iterations)\n",
          table(ev[:first + 1], "first CFG denoise step (graph replay)"),
          table(ev[first + 1:second + 1], "second CFG denoise step (graph replay; no iteration prologue)"),
          table(ev, f"whole iteration (k = {args.k} denoise steps + tail + optimizer)")]
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        f.write("\n".join(md))
    print("\n".join(md)[:9000])


if __name__ == "__main__":
    main()
