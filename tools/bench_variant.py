"""Dev tool: run bench.py against a library variant (tools/bin/libact_<name>.so).  usage: python tools/bench_variant.py <name> [bench args]"""
import importlib, os, runpy, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
act = importlib.import_module("anonymous-credit-tokens_b200")
act.LIB_PATH = os.path.join(ROOT, "tools", "bin", f"libact_{sys.argv[1]}.so")
sys.argv = [os.path.join(ROOT, "bench.py")] + sys.argv[2:]
runpy.run_path(sys.argv[0], run_name="__main__")
