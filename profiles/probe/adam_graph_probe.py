import torch, torch.nn as nn, torch.nn.functional as F
torch.manual_seed(0)
dev='cuda'
def run(fused, graph):
    torch.manual_seed(0)
    m = nn.Sequential(nn.Linear(8,64), nn.ReLU(), nn.Linear(64,1)).to(dev)
    opt = torch.optim.Adam(m.parameters(), lr=1e-3, capturable=True, fused=fused)
    x = torch.randn(256,8,device=dev); y = x.sum(1,keepdim=True)
    def step():
        loss = F.mse_loss(m(x), y)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3): step()
    torch.cuda.current_stream().wait_stream(s)
    if graph:
        opt.zero_grad(set_to_none=True)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g): step()
        for _ in range(500): g.replay()
    else:
        for _ in range(501): step()
    torch.cuda.synchronize()
    return float(F.mse_loss(m(x), y)), [p.detach().clone() for p in m.parameters()], opt.state_dict()['state'][0]['step']
for fused in (False, True):
    for graph in (False, True):
        l, p, st = run(fused, graph)
        print(fused, graph, l, float(p[0].abs().sum()), st)
