"""Same-box A/B of two builds of the CUDA library (or two environment settings) on the full joint step.

gpurun boxes differ by about +-4 % on the same binary, so a change must be judged on ONE box: this runs bench.py
alternately with each variant (A B A B ...) in one process tree and prints the paired step times.

  # two libraries: build the variant next to the default one (it travels with the snapshot because *.so is not gpurun-ignored)
  nvcc ... -o deepatlas_b200/lib_variant.so ...
  gpurun -- python tools/ab_bench.py --b-lib deepatlas_b200/lib_variant.so --rounds 2
  # two environment settings of the same library
  gpurun -- python tools/ab_bench.py --b-env DA_JOINT_UNFUSED=1

Numbers are bench.py's own (CUDA events, max over ranks); nothing here runs under a profiler."""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(env_extra, steps, warmup):
    env = dict(os.environ)
    env.update(env_extra)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", str(steps), "--warmup", str(warmup), "--no-cpu-baseline", "--no-torch-cuda", "--no-parity"],
                         env=env, capture_output=True, text=True, cwd=ROOT)
    for line in reversed(out.stdout.strip().splitlines()):
        if line.startswith("{"):
            return json.loads(line)["ms_per_step"]
    raise RuntimeError("bench.py printed no JSON line:\n" + out.stdout[-2000:] + out.stderr[-2000:])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--b-lib", default=None, help="path of the variant library (DA_LIB_PATH of arm B)")
    ap.add_argument("--b-env", action="append", default=[], help="KEY=VALUE set only for arm B (repeatable)")
    ap.add_argument("--rounds", type=int, default=2)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    b_env = dict(kv.split("=", 1) for kv in a.b_env)
    if a.b_lib:
        b_env["DA_LIB_PATH"] = os.path.abspath(a.b_lib)
    if not b_env:
        ap.error("nothing distinguishes arm B: give --b-lib and/or --b-env")
    A, B = [], []
    for r in range(a.rounds):
        A.append(run({}, a.steps, a.warmup))
        B.append(run(b_env, a.steps, a.warmup))
        print(f"round {r}: A {A[-1]:.3f} ms   B {B[-1]:.3f} ms   B - A {B[-1] - A[-1]:+.3f} ms", flush=True)
    ma, mb = sum(A) / len(A), sum(B) / len(B)
    print(json.dumps({"A_ms": A, "B_ms": B, "mean_A": ma, "mean_B": mb, "delta_ms": mb - ma, "delta_pct": 100.0 * (mb - ma) / ma, "B": b_env}))


if __name__ == "__main__":
    main()
