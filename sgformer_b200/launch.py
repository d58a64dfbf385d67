"""Run an UNMODIFIED reference training script with the B200 `ours.py` in front of its own:

    python -m sgformer_b200.launch --variant large /path/to/SGFormer/large/main-batch.py --method sgformer --dataset pokec ...

sys.path becomes [dropin/<variant>, <script dir>, ...] so `from ours import *` in the reference's parse.py picks the
drop-in while every other sibling import (logger, dataset, data_utils, eval, parse) stays the reference's."""
import argparse
import os
import runpy
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def main(argv=None):
    ap = argparse.ArgumentParser(prog="python -m sgformer_b200.launch")
    ap.add_argument("--variant", required=True, choices=["large", "medium", "100M"])
    ap.add_argument("--precision", choices=["bf16", "fp32"], default=None)
    ap.add_argument("script")
    ap.add_argument("script_args", nargs=argparse.REMAINDER)
    a = ap.parse_args(argv)
    if a.precision:
        os.environ["SGFORMER_B200_PRECISION"] = a.precision
    script = os.path.abspath(a.script)
    sdir = os.path.dirname(script)
    sys.path[:0] = [os.path.join(HERE, "dropin", a.variant), sdir]
    sys.argv = [script] + a.script_args
    os.chdir(sdir)
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
