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
    bead_diffusion/sources/{big_bead,small_bead,centre}_msd.csv (first 200 rows) -> bead_{...}_msd_head.csv
        examples/jfm.py:134-171 (np.savetxt of the ensemble mean-square displacements, 256 trajectories)
    rotational/sources/curve_{1,3,5}.csv -> rotational_curve_{1,3,5}.csv
        the three components of the rotation-vector variance of
        examples/published_problems/evensen_nalum_elgsaeter_macromol_2008.py (plot_correlations.py:7-13)
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
# ensemble mean-square displacements of the same study (256 trajectories; row k = lag of k snapshots = 2k time units)
COPIES += [(f"bead_diffusion/sources/{k}_msd.csv", f"bead_{k}_msd_head.csv", 200) for k in ("big_bead", "small_bead", "centre")]
# rotational Brownian motion (examples/published_problems/evensen_nalum_elgsaeter_macromol_2008.py): <delta u_k^2>(t)
COPIES += [(f"rotational/sources/curve_{k}.csv", f"rotational_curve_{k}.csv", None) for k in (1, 3, 5)]


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
