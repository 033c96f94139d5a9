"""Developer tool: same-box A/B of an environment knob -- boxes differ by up to 12 % in the power-capped clock, so an
effect is only meaningful as a pair measured inside ONE gpurun call.

    python tools/ab.py FASTA_B200_SPECULATE=1,0 [OTHER=a,b ...] [--repeat 2] -- python tools/tv_probe.py 4096 60
    python tools/ab.py FASTA_B200_AFFINE_PROBE=1,0 --json value,iters_per_sec_in_loop,roofline.avg_launch_ms -- \\
        python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e

Runs the command once per combination of the listed values (interleaved, `--repeat` rounds) in fresh processes and
prints the command's output lines prefixed with the setting; with --json only the named (dotted) keys of the
command's JSON lines.
"""
import itertools
import json
import os
import subprocess
import sys


def main(argv):
    if "--" not in argv:
        sys.exit(__doc__)
    head, cmd = argv[:argv.index("--")], argv[argv.index("--") + 1:]
    repeat, keys, knobs = 1, None, []
    it = iter(head)
    for a in it:
        if a == "--repeat":
            repeat = int(next(it))
        elif a == "--json":
            keys = next(it).split(",")
        else:
            name, vals = a.split("=", 1)
            knobs.append((name, vals.split(",")))
    for _ in range(repeat):
        for combo in itertools.product(*[v for _, v in knobs]):
            env = dict(os.environ)
            tag = " ".join(f"{n}={v}" for (n, _), v in zip(knobs, combo))
            env.update({n: v for (n, _), v in zip(knobs, combo)})
            out = subprocess.run(cmd, env=env, capture_output=True, text=True)
            if out.returncode != 0:
                print(f"[{tag}] FAILED rc={out.returncode}: {out.stderr[-400:]}", flush=True)
                continue
            for line in out.stdout.splitlines():
                if keys is None:
                    print(f"[{tag}] {line}", flush=True)
                elif line.startswith("{"):
                    d = json.loads(line)
                    vals = []
                    for k in keys:
                        v = d
                        for part in k.split("."):
                            v = v.get(part) if isinstance(v, dict) else None
                        vals.append(f"{k}={v:.4g}" if isinstance(v, float) else f"{k}={v}")
                    print(f"[{tag}] " + " ".join(vals), flush=True)


if __name__ == "__main__":
    main(sys.argv[1:])
