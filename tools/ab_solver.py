#!/usr/bin/env python
"""A/B timing of solver builds on the bench workload: the workload is generated once, then every
library given on the command line solves it in its own process (FEA_B200_LIB).

    python tools/ab_solver.py [--plates 100] [--reps 5] lib_a.so lib_b.so ...

Prints per library: solve ms of every repetition, iteration sum, status histogram and a checksum of
u (builds that only differ in synchronisation must print the same checksum)."""
import os, sys, json, pickle, subprocess, hashlib
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(path, reps):
    from fea_diffusion_b200 import Context, pack
    with open(path, "rb") as f:
        samples = pickle.load(f) * int(os.environ.get("AB_DUP", "1"))
    ctx = Context(0)
    packed = pack(samples, alloc=ctx.pinned_empty)
    ms = []
    with ctx.create_batch(packed) as b:
        b.assemble()
        for _ in range(reps):
            ctx.event_record(0)
            b.solve(1e-10, 20000)
            ctx.event_record(1)
            ctx.synchronize()
            ms.append(round(ctx.event_elapsed_ms(0, 1), 3))
        st = b.stats()
        r = b.download()
    print(json.dumps({"lib": os.path.basename(os.environ.get("FEA_B200_LIB", "default")), "solve_ms": ms,
                      "iters": int(r.iters.sum()), "status": np.bincount(r.status + 1).tolist(),
                      "u_sha": hashlib.sha256(np.ascontiguousarray(r.u).tobytes()).hexdigest()[:16],
                      "cluster_ms": st["cluster_ms"]}), flush=True)
    ctx.close()


def main():
    args = sys.argv[1:]
    plates, reps = 100, 5
    while args and args[0].startswith("--"):
        k = args.pop(0)
        v = int(args.pop(0))
        if k == "--plates": plates = v
        elif k == "--dup": os.environ["AB_DUP"] = str(v)   # solve the workload v times over in ONE batch
        elif k == "--reps": reps = v
    if args and args[0] == "__child__":
        return child(args[1], int(args[2]))
    from fea_diffusion_b200.workload import build_workload
    items, _ = build_workload(plates, 4, 64)
    path = "/tmp/ab_workload.pkl"
    with open(path, "wb") as f:
        pickle.dump([it.setup.sample for it in items], f)
    for lib in args or [""]:
        env = dict(os.environ)
        if lib:
            env["FEA_B200_LIB"] = os.path.abspath(lib)
        subprocess.run([sys.executable, os.path.abspath(__file__), "__child__", path, str(reps)], env=env)


if __name__ == "__main__":
    main()
