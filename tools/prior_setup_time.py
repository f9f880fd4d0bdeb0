"""Time of the one-off prior set-up product (train_insilico.py:208-209: batch_for_prior @ prior_mat, 10 000 x 11 165 rows against
a 3 % dense 11 165 x 11 165 prior) through phoenix_b200.prior_grad_from_matrix, and its error against float64."""
import torch, time, sys
sys.path.insert(0, '.')
import phoenix_b200 as pb
G, B = 11165, 10000
gen = torch.Generator().manual_seed(0)
pm = ((torch.rand(G, G, generator=gen) < 0.03).float() * (torch.rand(G, G, generator=gen) - 0.5)).cuda()
x = (torch.rand(B, 1, G, generator=gen) - 0.5).cuda()
pb.prior_grad_from_matrix(x[:8], pm)
torch.cuda.synchronize()
for _ in range(2):
    t0 = time.perf_counter(); out = pb.prior_grad_from_matrix(x, pm); torch.cuda.synchronize(); t1 = time.perf_counter()
    print("prior_grad_from_matrix 10000 x 11165, 3%% prior: %.1f ms (conversion to CSC included)" % ((t1 - t0) * 1e3))
ref = (x.view(B, G)[:64].double() @ pm.double())
print("rel err vs float64 (64 rows): %.2e" % float((out.view(B, G)[:64].double() - ref).norm() / ref.norm()))
