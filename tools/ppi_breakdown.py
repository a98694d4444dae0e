"""One eager PPI epoch (benchmarks.epochs) for an ncu launch list: which kernels the 3-layer step spends its time in."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from benchmarks import epochs
ms, info = epochs.ppi_epoch_ms(torch.device("cuda", 0), epochs=1, warmup=1)
print("ppi epoch ms", ms, info)
