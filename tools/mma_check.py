"""Development aid: the small-n tensor-core kernel (sde_kernel_mma.cuh) against the oracle (production noise
through ccvm_dump_noise) and against the tiled kernel's timing.  usage: mma_check.py [--quick] [--n 70]"""
import argparse
import json
import os
import sys
import time
import traceback

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["CCVM_MMA"] = "1"

import torch  # noqa: E402

from tests.test_gpu_parity import parity_case  # noqa: E402
from tests.test_gpu_production_parity import TILES, launch_info, tol_of  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, nargs="*", default=[70])
    ap.add_argument("--only", default="")
    ap.add_argument("--batch", type=int, default=200)
    ap.add_argument("--iters", type=int, default=150)
    args = ap.parse_args()
    bad = 0
    for n in args.n:
        for solver, adam in TILES:
            name = solver + ("_adam" if adam else "")
            if args.only and name not in args.only.split(","):
                continue
            try:
                info = launch_info(solver, adam, n, args.batch, args.iters)
                assert info["threads"] in (288, 544), info
                t0 = time.time()
                err = parity_case(solver, adam, n, args.batch, args.iters, tol_of(solver, adam), philox=(77, 5 * n))
                print(json.dumps({"n": n, "tile": name, "batch": args.batch, "iters": args.iters, "rel_obj_err": err,
                                  "ok": True, "regs": info["regs"], "s": round(time.time() - t0, 2)}), flush=True)
            except Exception as e:  # noqa: BLE001
                bad += 1
                print(json.dumps({"n": n, "tile": name, "ok": False, "error": str(e)[:300]}), flush=True)
                if "CUDA" in str(e) or "cuda" in str(e):
                    traceback.print_exc()
                    return 2
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
