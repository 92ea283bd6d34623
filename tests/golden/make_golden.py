"""Provenance of the fixtures in this directory.

They are DATA files the reference ships with its paper scripts (not source code), copied byte for byte - the two
bead files cut to their first 200 rows - because /root/reference does not exist on the GPU box.  Run this script in a
container that has the reference mounted to re-create (or verify, with --check) them.

    reference path (under examples/examples_from_paper/paper_graphics_python/)  ->  fixture
    stopping_time/sources/trajectory_{1..5}.csv           -> stopping_time_trajectory_{1..5}.csv
        output of the reference's hitting-time study (examples/hit.py / examples_from_paper/
        polar_random_walk_hitting_time.py; read by stopping_time/plot_trajectories.py:7-11): Euler, fp32 - the recipe
        (problem, seeds, dt, columns) that reproduces them is in
        tests/test_oracle_golden.py::test_golden_stopping_time_trajectories
    stopping_time/sources/{strong,weak}_convergence_test.csv -> same names (error-vs-dt tables of the paper figure)
    bead_diffusion/sources/single_trajectory.csv (first 200 rows) -> bead_single_trajectory_head.csv
    bead_diffusion/sources/time.csv (first 200 rows)              -> bead_time_head.csv
        output of examples/jfm.py (four RPY beads, Euler, fp32; read by bead_diffusion/
        plot_mean_square_displacement.py:10); recipe in test_golden_bead_trajectory
"""
import filecmp
import os
import shutil
import sys

REF = os.environ.get("REFERENCE_ROOT", "/root/reference")
SRC = os.path.join(REF, "examples", "examples_from_paper", "paper_graphics_python")
HERE = os.path.dirname(os.path.abspath(__file__))

COPIES = [(f"stopping_time/sources/trajectory_{i}.csv", f"stopping_time_trajectory_{i}.csv", None) for i in range(1, 6)]
COPIES += [(f"stopping_time/sources/{k}_convergence_test.csv", f"{k}_convergence_test.csv", None) for k in ("strong", "weak")]
COPIES += [("bead_diffusion/sources/single_trajectory.csv", "bead_single_trajectory_head.csv", 200),
           ("bead_diffusion/sources/time.csv", "bead_time_head.csv", 200)]


def main(check):
    bad = 0
    for src, dst, head in COPIES:
        src, dst = os.path.join(SRC, src), os.path.join(HERE, dst)
        if head is None:
            if check:
                same = filecmp.cmp(src, dst, shallow=False)
            else:
                shutil.copyfile(src, dst)
                same = True
        else:
            with open(src) as f:
                lines = f.readlines()[:head]
            if check:
                with open(dst) as f:
                    same = f.readlines() == lines
            else:
                with open(dst, "w") as f:
                    f.writelines(lines)
                same = True
        print(("ok   " if same else "DIFF ") + os.path.basename(dst))
        bad += not same
    return bad


if __name__ == "__main__":
    sys.exit(main("--check" in sys.argv))
