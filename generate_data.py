#!/usr/bin/env python
"""CLI with the reference's flags (reference generate_data.py:5-44) over the B200 path.

    python generate_data.py --num_plates 100 --conditions_per_plate 4 --steps_per_condition 11 \\
        --image_size 64 --save_displacement --data_dir data [--gpus 8] [--backend batched|dropin]

`--backend batched` (default) runs the pipelined, sharded generator (fea_diffusion_b200.dataset) --
one process per GPU, plates dealt round-robin, no communication.  `--backend dropin` runs the
reference's OWN sequential loop, unchanged, over the CUDA-backed FEAnalysis: it needs a checkout of
the reference (`--reference_dir` or FEA_REFERENCE_DIR), whose datagen/generate.py is loaded with its
imports bound to the drop-in classes (fea_diffusion_b200.datagen.generate.load_reference_generate).
BASELINE.json spells two flags differently from the reference (`--num_conditions_per_plate`,
`--num_steps`): both spellings are accepted.  `--mesh_size` is a float here (the reference declares
it `type=int`, which makes every value but the default unusable)."""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse(argv=None):
    ap = argparse.ArgumentParser(description="Generate data for training.")
    ap.add_argument("--num_plates", type=int, default=1, help="Number of plates to generate.")
    ap.add_argument("--start_plate", type=int, default=None, help="Plate index to start generating from.")
    ap.add_argument("--conditions_per_plate", "--num_conditions_per_plate", dest="conditions_per_plate", type=int,
                    default=4, help="Number of conditions to generate per plate.")
    ap.add_argument("--steps_per_condition", "--num_steps", dest="steps_per_condition", type=int, default=11,
                    help="Number of steps to generate per condition.")
    ap.add_argument("--mesh_size", type=float, default=1e-2, help="Mesh size.")
    ap.add_argument("--image_size", type=int, default=512, help="Image size.")
    ap.add_argument("--save_meshes", action="store_true", help="Save meshes per condition.")
    ap.add_argument("--save_displacement", action="store_true", help="Save displacement images.")
    ap.add_argument("--save_strain", action="store_true", help="Save strain images.")
    ap.add_argument("--save_stress", action="store_true", help="Save stress images.")
    ap.add_argument("--data_dir", type=str, default="data", help="Data directory.")
    ap.add_argument("--use_wandb", action="store_true", help="Use wandb.")
    ap.add_argument("--wandb_project", type=str, help="Wandb project name.")
    ap.add_argument("--wandb_restrict_cache", type=int, default=10, help="Restrict wandb cache.")
    ap.add_argument("--backend", choices=["batched", "dropin"], default="batched")
    ap.add_argument("--reference_dir", type=str, default=os.environ.get("FEA_REFERENCE_DIR"),
                    help="checkout of namanxkumar/fea-diffusion for --backend dropin")
    ap.add_argument("--gpus", type=int, default=1, help="GPUs of this box to shard the plates over (batched).")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--plates_per_batch", type=int, default=50)
    ap.add_argument("--region_method", choices=["reference", "lloyd"], default="reference",
                    help="material regions of the condition sampler: the reference's KMeans / agglomerative methods "
                         "(scikit-learn) or the fast Lloyd stand-in of the benchmark workloads (batched backend)")
    ap.add_argument("--rank", type=int, default=None, help=argparse.SUPPRESS)   # set for the per-GPU children
    return ap.parse_args(argv)


def main(argv=None):
    a = parse(argv)
    assert a.save_displacement or a.save_strain or a.save_stress, \
        "Must save at least one of displacement, strain, or stress."
    import __graft_entry__ as ge
    ge.build()
    if a.backend == "dropin":
        if not a.reference_dir:
            sys.exit("--backend dropin runs the reference's own datagen/generate.py: pass --reference_dir")
        from fea_diffusion_b200.datagen.generate import load_reference_generate
        load_reference_generate(a.reference_dir)(
            data_dir=a.data_dir, image_size=a.image_size, num_plates=a.num_plates, start_plate=a.start_plate,
            conditions_per_plate=a.conditions_per_plate, mesh_size=a.mesh_size, save_displacement=a.save_displacement,
            save_strain=a.save_strain, save_stress=a.save_stress, num_steps_per_condition=a.steps_per_condition,
            save_meshes=a.save_meshes)
        return
    start = max(0, (a.start_plate - 1) if a.start_plate is not None else 0)   # reference generate.py:50 quirk
    if a.gpus > 1 and a.rank is None:
        procs = [subprocess.Popen([sys.executable, os.path.abspath(__file__)] + sys.argv[1:] + ["--rank", str(r)])
                 for r in range(a.gpus)]
        codes = [p.wait() for p in procs]          # negative = killed by a signal (e.g. -9: out of memory)
        for r, rc in enumerate(codes):
            if rc != 0:
                print("rank %d failed with exit code %d: its plates are missing from %s" % (r, rc, a.data_dir), file=sys.stderr)
        sys.exit(1 if any(rc != 0 for rc in codes) else 0)
    from fea_diffusion_b200.dataset import generate_dataset
    rank = a.rank or 0
    st = generate_dataset(a.data_dir, a.num_plates, a.conditions_per_plate, a.image_size, a.steps_per_condition,
                          a.mesh_size, seed=a.seed, rank=rank, world=a.gpus, plates_per_batch=a.plates_per_batch,
                          save_meshes=a.save_meshes, start_plate=start, save_displacement=a.save_displacement,
                          save_stress=a.save_stress, save_strain=a.save_strain, region_method=a.region_method,
                          progress=lambda d, n: print("rank %d: %d / %d plates" % (rank, d, n), flush=True))
    print(json.dumps(st))


if __name__ == "__main__":
    main()
