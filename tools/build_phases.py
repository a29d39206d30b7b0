import os, sys, time
sys.path.insert(0, os.getcwd())
import torch, csfm_b200 as fm
dev = torch.device("cuda", 0)
for name, n, kind in (("c3", 1 << 30, "byte"), ("c5", 4_000_000_000, "dna")):
    text = (fm.workloads.byte_text_torch if kind == "byte" else fm.workloads.dna_text_torch)(n, 3, dev)
    torch.cuda.synchronize(); torch.cuda.empty_cache()
    for rep in range(2):
        t0 = time.perf_counter()
        idx = fm.FMIndex.build_from_text_device(text.data_ptr(), n, fm.BuildParams(ssa_stride=32), device=0)
        torch.cuda.synchronize()
        print(f"== {name} rep {rep}: {time.perf_counter() - t0:.3f} s", file=sys.stderr, flush=True)
        idx.close()
    del text
    torch.cuda.empty_cache()
