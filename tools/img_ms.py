import sys; sys.path.insert(0, "/root/repo")
import torch, bench
dev = torch.device("cuda:0")
for i in range(3):
    print(bench.image_configs_ms(dev, "f16x2", False))
