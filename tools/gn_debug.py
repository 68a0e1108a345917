import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200sd import ops
DEV='cuda'
torch.manual_seed(1)
for (B,hw,C0,C1,silu) in [(4,16,128,0,True),(4,16,128,128,True),(4,64,128,0,False),(4,1024,64,0,True),(4,256,128,64,True),(4,64,128,128,True)]:
    C=C0+C1
    x0=(torch.randn(B*hw,C0,device=DEV)*1.5+0.3); x1=torch.randn(B*hw,C1,device=DEV) if C1 else None
    gamma=torch.randn(C,device=DEV)*0.5+1; beta=torch.randn(C,device=DEV)*0.2
    dy=torch.randn(B*hw,C,device=DEV).bfloat16()
    t=torch.empty(B*hw,C,device=DEV,dtype=torch.bfloat16); st=torch.empty(B,32,2,device=DEV)
    ops.groupnorm_silu(x0,x1,gamma,beta,t,B,hw,32,1e-5,silu,stats_out=st)
    outs=[]
    for name,mr in (("recompute",None),("fast",st),("fast",st)):
        o0=torch.zeros(B*hw,C0,device=DEV); o1=torch.zeros(B*hw,C1,device=DEV) if C1 else None
        dg=torch.zeros(C,device=DEV); db=torch.zeros(C,device=DEV)
        ops.groupnorm_silu_bwd(x0,x1,gamma,beta,dy,o0,o1,B,hw,dgamma=dg,dbeta=db,eps=1e-5,silu=silu,mean_rstd=mr)
        torch.cuda.synchronize()
        outs.append((o0,dg,db))
    r,f1,f2=outs
    print((B,hw,C0,C1,silu), "recompute-vs-fast dx %.2e dg %.2e db %.2e | fast-vs-fast dx %.2e" % (
        float((r[0]-f1[0]).abs().max()/r[0].abs().max()), float((r[1]-f1[1]).abs().max()/r[1].abs().max()), float((r[2]-f1[2]).abs().max()/r[2].abs().max()),
        float((f1[0]-f2[0]).abs().max()/r[0].abs().max())))
