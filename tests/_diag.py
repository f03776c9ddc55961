import os, sys
sys.path.insert(0, os.getcwd())
os.environ["KMCUDA_B200_DEBUG"]="1"
import numpy as np, torch
from kmcuda_b200.shard import Shard
for (n,d,k) in [(1024,64,64),(4096,256,1024)]:
    X=torch.rand((n,d),device="cuda"); C=torch.rand((k,d),device="cuda")
    sh=Shard(n,d,k)
    a=torch.full((n,),-1,dtype=torch.int32,device="cuda"); p=a.clone(); ch=torch.zeros(1,dtype=torch.int32,device="cuda")
    try:
        sh.assign(X,C,a,p,ch); torch.cuda.synchronize(); print("ok",n,d,k,hex(sh.last_error()), sh.last_pass_info())
    except Exception as e:
        print("fail",n,d,k,e)
