import sys, os
sys.path.insert(0, os.getcwd())
import torch, spectrum_analyzer_b200 as sa
from spectrum_analyzer_b200 import cabi
n=int(sys.argv[1]); frames=int(sys.argv[2]); sc=int(sys.argv[3]); win=int(sys.argv[4]); st=int(sys.argv[5])
lib=cabi.load(); dev=torch.device("cuda",0); stream=torch.cuda.current_stream(dev)
ctx=sa.Context(0, stream.cuda_stream); chk=sa.api._check
x=torch.empty(frames*n,dtype=torch.float32,device=dev); chk(lib.sa_generate_noise(ctx._h,x.data_ptr(),0,x.numel(),1))
nb=n//2+1; vals=torch.empty(frames*nb,dtype=torch.float32,device=dev); stats=torch.empty(frames*32,dtype=torch.uint8,device=dev)
for _ in range(4):
    chk(lib.sa_spectrum_batched(ctx._h,x.data_ptr(),frames,n,n,44100,0,0.0,0.0,win,sc,st,vals.data_ptr(),nb,stats.data_ptr(),None,None))
torch.cuda.synchronize(); print("ok")
