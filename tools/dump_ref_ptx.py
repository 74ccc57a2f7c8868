import sys, os
sys.path.insert(0, "/root/repo")
from oracle import ref_ocl
assert ref_ocl.available(), ref_ocl.why_unavailable()
for f, e in (("v210.cl", "read"), ("transform.cl", "transform"), ("combine_2.cl", "combine_2")):
    t = ref_ocl.program_text(f, e)
    open(f"/root/repo/gpurun_out/ref_{e}.ptx", "w").write(t)
    print(f, e, len(t))
