"""Which summation order does torch use for the 4-element L2 norm behind F.normalize on CUDA? (dev probe)"""
import torch
torch.manual_seed(0)
q = torch.randn(2_000_000, 4, device="cuda")
n = torch.linalg.vector_norm(q, 2, dim=1)
x = [q[:, i].contiguous() for i in range(4)]
sq = [xi * xi for xi in x]
cands = {
    "seq ((0+1)+2)+3": torch.sqrt(((sq[0] + sq[1]) + sq[2]) + sq[3]),
    "pair (0+1)+(2+3)": torch.sqrt((sq[0] + sq[1]) + (sq[2] + sq[3])),
    "pair (0+2)+(1+3)": torch.sqrt((sq[0] + sq[2]) + (sq[1] + sq[3])),
    "fma chain": torch.sqrt(torch.addcmul(torch.addcmul(torch.addcmul(sq[0], x[1], x[1]), x[2], x[2]), x[3], x[3])),
    "fma chain rev": torch.sqrt(torch.addcmul(torch.addcmul(torch.addcmul(sq[3], x[2], x[2]), x[1], x[1]), x[0], x[0])),
}
for k, v in cands.items():
    print(f"{k:22s} mismatches: {int((v != n).sum())}")
qn = torch.nn.functional.normalize(q)
print("normalize == q / n.clamp_min(1e-12):", bool(torch.equal(qn, q / n.clamp_min(1e-12).unsqueeze(1))))
o = torch.randn(1_000_000, device="cuda") * 3
print("sigmoid == 1/(1+exp(-x)):", int((torch.sigmoid(o) != 1.0 / (1.0 + torch.exp(-o))).sum()))
