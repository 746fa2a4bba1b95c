import torch, json
out={}
for (M,N,K) in [(1<<18,64,64),(1<<17,64,64),(1<<18,64,32),(4096,4096,4096),(1<<13,1<<13,1<<11)]:
    for dt,name in ((torch.complex128,"z"),(torch.complex64,"c")):
        a=torch.randn(M,K,dtype=dt,device="cuda"); b=torch.randn(K,N,dtype=dt,device="cuda")
        # column-major semantics: C^T = B^T A^T; use row-major equivalents (same flops)
        for _ in range(3): torch.matmul(a,b)
        torch.cuda.synchronize()
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        best=1e9
        for _ in range(5):
            e0.record(); torch.matmul(a,b); e1.record(); torch.cuda.synchronize()
            best=min(best,e0.elapsed_time(e1))
        out["%sgemm_%dx%dx%d"%(name,M,N,K)]=round(8.0*M*N*K/(best*1e-3)/1e12,2)
        # transposed operand variant (A stored K x M)
        at=torch.randn(K,M,dtype=dt,device="cuda")
        for _ in range(3): torch.matmul(at.t(),b)
        torch.cuda.synchronize(); best=1e9
        for _ in range(5):
            e0.record(); torch.matmul(at.t(),b); e1.record(); torch.cuda.synchronize()
            best=min(best,e0.elapsed_time(e1))
        out["%sgemm_T_%dx%dx%d"%(name,M,N,K)]=round(8.0*M*N*K/(best*1e-3)/1e12,2)
print(json.dumps(out,indent=1))
