#!/usr/bin/env python
"""Run one of the reference's launcher scripts (pointwise.sh, reward_pair_dataloader.sh, ppo.sh, ppo_eval.sh) against
the drop-in tree on N GPUs of this box.

The `.sh` files hard-code `CUDA_VISIBLE_DEVICES=0,1,2,3 torchrun --nproc_per_node=4` (ppo.sh:59); everything else in
them is the flag arrays.  This launcher reads the UNCHANGED script -- variable assignments, the `name=( ... )` arrays,
the `mkdir -p` lines and the final torchrun line naming `finetune/<stage>.py` and the array order -- and starts

    python -m torch.distributed.run --nproc-per-node N ... dropin/finetune/<stage>.py <the same flags> [extra flags]

in the current working directory (which must hold LRMovieNet/ and models/ like the reference checkout).

    python dropin/launch.py /path/to/ppo.sh my_experiment --gpus 8
    python dropin/launch.py ppo.sh exp --gpus 1 -- --update_timesteps 4 --batch_size 2      # flags after -- win
"""
import argparse
import os
import re
import shlex
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def parse_script(path, exp_name):
    """-> (stage file e.g. 'finetune/ppo.py', flag list, directories to create)."""
    text = open(path).read()
    env = {"1": exp_name}

    def expand(s):
        return re.sub(r"\$\{?(\w+)\}?", lambda m: env.get(m.group(1), os.environ.get(m.group(1), "")), s)

    arrays, mkdirs, stage, order = {}, [], None, []
    lines = iter(text.splitlines())
    for raw in lines:
        line = raw.split("#", 1)[0].rstrip() if not raw.lstrip().startswith("#") else ""
        if not line.strip():
            continue
        m = re.match(r"^\s*(\w+)=\((.*)$", line)
        if m:                                                # array: collect until the closing parenthesis
            name, body = m.group(1), m.group(2)
            while ")" not in body:
                nxt = next(lines)
                body += " " + nxt.split("#", 1)[0]
            arrays[name] = [expand(tok) for tok in shlex.split(body[:body.index(")")])]
            continue
        m = re.match(r"^\s*(\w+)=(\S*)\s*$", line)
        if m:
            env[m.group(1)] = expand(m.group(2))
            continue
        m = re.match(r"^\s*mkdir\s+-p\s+(\S+)", line)
        if m:
            mkdirs.append(expand(m.group(1)))
            continue
        if "torchrun" in line or stage is not None:
            m = re.search(r"(finetune/\w+\.py)", line)
            if m:
                stage = m.group(1)
            order += re.findall(r"\$\{(\w+)\[@\]\}", line)
    if stage is None:
        raise SystemExit(f"{path}: no `torchrun ... finetune/<stage>.py` line found")
    flags = [tok for name in order for tok in arrays[name]]
    return stage, flags, mkdirs


def main():
    argv = sys.argv[1:]
    extra = []
    if "--" in argv:
        k = argv.index("--")
        argv, extra = argv[:k], argv[k + 1:]
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("script", help="one of the reference's launcher .sh files (read, not executed)")
    ap.add_argument("exp_name", help="the positional $1 of the script")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--master-port", type=int, default=29576)
    ap.add_argument("--dry-run", action="store_true", help="print the command and exit")
    args = ap.parse_args(argv)
    stage, flags, mkdirs = parse_script(args.script, args.exp_name)
    # later occurrences win in argparse, so overrides are simply appended
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
           "--master-addr", "127.0.0.1", "--master-port", str(args.master_port), os.path.join(HERE, stage)] + flags + extra
    if args.dry_run:
        print(" ".join(shlex.quote(c) for c in cmd))
        return 0
    for d in mkdirs:
        os.makedirs(d, exist_ok=True)
    return subprocess.call(cmd)


if __name__ == "__main__":
    sys.exit(main())
