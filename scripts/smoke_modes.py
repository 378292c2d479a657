import sys, torch
sys.path.insert(0, ".")
import __graft_entry__ as g
mode = sys.argv[1]
if mode == "det":
    torch.use_deterministic_algorithms(True, warn_only=True)
g.smoke()
