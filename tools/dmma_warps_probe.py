import numpy as np, sys
sys.path.insert(0,'/root/repo')
import picoquant_jl_b200
from picoquant_jl_b200.host.b200_backend import B200Backend
b=B200Backend(np.complex128)
for w in (4,8,12,16,24,32):
    print(w, "warps/SM:", round(b.microbench("dmma_tflops_w%d"%w),2), "TF")
print("default:", b.microbench("dmma_tflops"))
