import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["SWIN_ONLY"] = os.environ.get("SWIN_ONLY", "cfg4")
from long_context_biomedical_imaging_b200 import ops
ops.set_window_kernel_mode(sys.argv[1])
exec(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "bench_swin.py")).read())
