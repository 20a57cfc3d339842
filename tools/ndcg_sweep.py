"""BASELINE config 5 (configs[4]): NDCG@k throughput sweep.  The sweep times the CPU oracle beside the GPU kernel, so it is
part of bench.py (the only non-test code allowed to execute oracle/):   python bench.py --ndcg-sweep OUT.md"""
import os
import subprocess
import sys

if __name__ == "__main__":
    out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/ndcg_sweep.md"
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.exit(subprocess.call([sys.executable, os.path.join(root, "bench.py"), "--ndcg-sweep", out]))
