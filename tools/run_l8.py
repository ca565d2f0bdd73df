import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import meta_viterbinet_b200 as mvn
dev = torch.device('cuda', 0)
L = int(sys.argv[1]) if len(sys.argv) > 1 else 8
torch.manual_seed(L)
net = torch.nn.Sequential(torch.nn.Linear(1, 100), torch.nn.Sigmoid(), torch.nn.Linear(100, 50), torch.nn.ReLU(), torch.nn.Linear(50, 2 ** L))
w = [p.detach().to(dev).contiguous() for p in net.parameters()]
y = torch.randn(1 << 16, 120, device=dev) * 1.5
for _ in range(3):
    out = mvn.ops.vnet_decode(y, w)
torch.cuda.synchronize()
print('ok')
