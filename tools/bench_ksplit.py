import sys, runpy
sys.path.insert(0, "/root/repo")
from asr_b200 import _lib
_lib.query("asrb_debug_rnn_ksplit", int(sys.argv[1]))
sys.argv = ["bench.py", "--steps", "10", "--warmup", "3", "--no-cpu-baseline"]
runpy.run_path("/root/repo/bench.py", run_name="__main__")
