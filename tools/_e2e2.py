import sys, os
sys.argv = ["x", "1000"]
import torch
torch.cuda.init(); x = torch.zeros(1, device="cuda")
exec(open("tools/e2e_timing.py").read())
