#!/usr/bin/env python3
"""Same-box A/B of library builds (boxes differ by +-3 %, so only same-call comparisons decide anything):
     python tools/ab.py [--rounds 2] [--secs 1.5] name=path.so ...        (path "-" = the in-tree library)
Each build is timed in its own process (the library is loaded once per process), rounds interleaved A B A B; for each
sparsity 0 / 42 / 77 % (balanced lists) and a Bernoulli 42 % list the forward runs back to back for --secs seconds at
the Wan2.1-14B shape and the mean ms of the second half is reported."""
import argparse, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child(secs):
    import torch
    sys.path.insert(0, ROOT)
    from liteattention_b200 import _native as N, synth
    B, S, H, D = 1, int(os.environ.get("S", 75600)), int(os.environ.get("H", 40)), 128
    g = torch.Generator(device="cuda").manual_seed(0)
    q, k, v = (torch.randn(B, S, H, D, device="cuda", generator=g).to(torch.bfloat16) for _ in range(3))
    out = torch.empty_like(q); lse = torch.empty(B, H, S, device="cuda")
    qt, kt = synth.tile_counts(S)
    stat = torch.empty(B, H, qt, kt, device="cuda")
    res = {}
    cases = [("0", None), ("42", synth.exact_sparsity_list(B, H, qt, kt, 0.42, seed=1234, device="cuda")[0]),
             ("77", synth.exact_sparsity_list(B, H, qt, kt, 0.77, seed=1234, device="cuda")[0]),
             ("42b", synth.random_skip_list(B, H, qt, kt, 0.42, seed=99, device="cuda")[0])]
    for name, rl in cases:
        st = stat if rl is not None else None
        for _ in range(2): N.fwd(q, k, v, out, lse, D ** -0.5, rl, st)
        torch.cuda.synchronize()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record(); N.fwd(q, k, v, out, lse, D ** -0.5, rl, st); e[1].record(); torch.cuda.synchronize()
        n = max(4, int(secs * 1e3 / e[0].elapsed_time(e[1])))
        e[0].record()
        for _ in range(n // 2): N.fwd(q, k, v, out, lse, D ** -0.5, rl, st)
        e[1].record()
        for _ in range(n // 2): N.fwd(q, k, v, out, lse, D ** -0.5, rl, st)
        e[2].record(); torch.cuda.synchronize()
        res[name] = e[1].elapsed_time(e[2]) / (n // 2)
    res["checksum"] = float(out.float().abs().mean())
    print("AB_RESULT " + json.dumps(res))


if __name__ == "__main__":
    if os.environ.get("AB_CHILD"):
        child(float(os.environ["AB_CHILD"]))
        sys.exit(0)
    ap = argparse.ArgumentParser()
    ap.add_argument("--rounds", type=int, default=2)
    ap.add_argument("--secs", type=float, default=1.5)
    ap.add_argument("libs", nargs="+")
    a = ap.parse_args()
    table = {}
    for r in range(a.rounds):
        for spec in a.libs:
            name, _, path = spec.partition("=")
            env = dict(os.environ, AB_CHILD=str(a.secs))
            if path and path != "-":
                env["LITEATTN_B200_LIB"] = os.path.abspath(path)
            p = subprocess.run([sys.executable, os.path.abspath(__file__)], env=env, capture_output=True, text=True, timeout=600)
            line = [l for l in p.stdout.splitlines() if l.startswith("AB_RESULT ")]
            if not line:
                print(f"{name}: FAILED rc={p.returncode} {p.stderr[-300:]}")
                continue
            res = json.loads(line[0][10:])
            table.setdefault(name, []).append(res)
            print(f"round {r} {name:12s} " + "  ".join(f"{k}={v:.3f}" for k, v in res.items()), flush=True)
    print("---- means")
    for name, rs in table.items():
        keys = [k for k in rs[0] if k != "checksum"]
        print(f"{name:12s} " + "  ".join(f"{k}={sum(r[k] for r in rs) / len(rs):.3f}" for k in keys))
