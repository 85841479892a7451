import sys; sys.path.insert(0,'/root/repo')
import torch
from cellcomm_b200 import ops
from tools.gemm_bench import timeit
M,N,K=128,3369,33694
a=ops.alloc2d(M,K); a.normal_(); b=ops.alloc2d(K,N); b.normal_(std=0.01)
bias=torch.zeros(N,device='cuda'); o=ops.alloc2d(M,N)
flush=torch.empty(256*1024*1024,dtype=torch.uint8,device='cuda')
for sp in (0,1,2,4,8,11,16,22,32):
    t=timeit(lambda: ops.gemm(M,N,[a],[b],[K],0,1,bias=bias,act=1,out16=o,splits=sp), flush=flush)
    print('splits',sp,'ms',round(t,4),'GB/s',round(2*K*N/t/1e6,1), flush=True)
