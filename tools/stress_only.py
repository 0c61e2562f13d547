import sys, json
sys.path.insert(0, '/root/repo')
sys.argv = sys.argv[:1]
import torch, bench
dev = torch.device('cuda:0')
print(json.dumps(bench.stress_leg(torch, dev, 6534.8), indent=1))
