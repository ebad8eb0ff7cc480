"""Dev tool: render one configuration a few times (for ncu): python tools/gpu_one.py VIEW ALG [N_ITER] [W H] [--noref]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from gpu_probe import run
from fractalshark_b200 import RenderAlgorithm as A
args = [a for a in sys.argv[1:] if not a.startswith("--")]
view = int(args[0]); alg = getattr(A, args[1])
n_iter = int(args[2]) if len(args) > 2 and args[2] != "0" else None
w, h = (int(args[3]), int(args[4])) if len(args) > 4 else (3840, 2160)
run(view, w, h, alg, n_iter, with_ref="--noref" not in sys.argv)
