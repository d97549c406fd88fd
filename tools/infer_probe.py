import sys, os
sys.path.insert(0, os.getcwd())
import torch, bench
torch.cuda.set_device(0)
print(bench.run_infer(torch.device("cuda", 0), 2, 3))
