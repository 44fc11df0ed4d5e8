import torch


def case(seed=0, N=2, P1=300, P2=411):
    g = torch.Generator().manual_seed(seed)
    a = torch.randn(N, P1, 3, generator=g) * 10
    b = torch.randn(N, P2, 3, generator=g) * 10
    b[0, 5] = b[0, 17]                     # duplicate points: ties resolve to the smaller index
    a[1, 3] = b[1, 100]                    # exact hit: distance 0
    return dict(a=a, b=b, la=torch.tensor([P1, P1 - 40]), lb=torch.tensor([P2 - 11, P2]),
                g=torch.randn(N, P1, 1, generator=g))
