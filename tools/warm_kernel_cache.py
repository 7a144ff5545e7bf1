"""Warm the in-tree cubin cache before a gpurun call: run every gpu-marked test on this (GPU-less) machine with newPsydiff
patched to compile the model at once; each test then stops at the first device call.  The cubins travel with the snapshot, so
the GPU box does not spend its minutes in nvcc.  (Models a test builds after its first device call are compiled on the box.)"""
import sys, os, time, inspect, importlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))

import psydiff_b200 as pd
from psydiff_b200 import model as M
seen = {}
orig_new = M.newPsydiff
def new(*a, **k):
    m = orig_new(*a, **k)
    t = time.time(); p = M.compileModelSource(m); dt = time.time() - t
    if dt > 0.05: print(f"   compiled {os.path.basename(p)} n={m['nlatent']} {dt:.1f}s", flush=True)
    return m
M.newPsydiff = new; pd.newPsydiff = new
for modname in ("test_gpu_parity", "test_gpu_examples", "test_gpu_optim", "test_r_shim"):
    mod = importlib.import_module(modname)
    for name, fn in inspect.getmembers(mod, inspect.isfunction):
        if not name.startswith("test_"): continue
        marks = getattr(fn, "pytestmark", [])
        params = [mk for mk in marks if mk.name == "parametrize"]
        argsets = [()]
        if params:
            argsets = [(v,) for v in params[0].args[1]]
        for a in argsets:
            try:
                if modname == "test_r_shim":
                    if name != "test_compile_fit_and_gradients_through_the_shim_on_the_gpu": continue
                    fn(mod.R())
                else:
                    fn(*a)
            except BaseException as e:
                print(name, a, "->", type(e).__name__, str(e)[:60].replace("\n"," "), flush=True)
