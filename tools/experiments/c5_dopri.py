import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import phoenix_b200 as pb
net = pb.ODENet("cuda", 20000, neurons=200)
y0 = torch.rand(4096, 20000, device="cuda")
with torch.no_grad():
    pb.odeint(net, y0, torch.tensor([0.0, 0.1]), method="dopri5", rtol=1e-5, atol=1e-7)
    torch.cuda.synchronize()
    pb.odeint(net, y0, torch.tensor([0.0, 0.1]), method="dopri5", rtol=1e-5, atol=1e-7)
    torch.cuda.synchronize()
