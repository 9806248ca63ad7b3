#!/usr/bin/env python
"""Build an INSTRUMENTED variant of the fused attention kernel (clock64 per phase of one softmax warp) into
tools/exp/libvar/libnhmc_attnprof.so: the source is generated from csrc/attn.cu by textual insertion, so there is no second
copy of the kernel to keep in step.  Then:  NHMC_LIB_PATH=$PWD/tools/exp/libvar/libnhmc_attnprof.so python tools/exp/attn_prof.py

    python tools/exp/make_attn_prof.py [--pad BYTES]      # --pad: extra shared memory per CTA (20000 forces one CTA per SM)
"""
import argparse
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CSRC = os.path.join(ROOT, "noise-space-hmc_b200", "csrc")


def instrument(s, pad):
    def rep(old, new):
        nonlocal s
        assert old in s, old
        s = s.replace(old, new, 1)
    rep('#include "tc5.cuh"', '#include "tc5.cuh"\n#define PROF_T() clock64()\n__device__ long long g_attn_prof[1024][12];\n')
    rep("static constexpr int TOTAL = OFF_BAR + 256;", "static constexpr int TOTAL = OFF_BAR + 256 + %d;" % pad)
    rep("    float m_used = -INFINITY;",
        "    long long pws = 0, pld = 0, pwpv = 0, prs = 0, nrs = 0, pwo = 0, pmx = 0, tq; const long long pstart = PROF_T();\n    float m_used = -INFINITY;")
    rep("      mbar_wait(bars + B_SFULL + hf, j & 1);\n      tc_fence_after();",
        "      tq = PROF_T(); mbar_wait(bars + B_SFULL + hf, j & 1); pws += PROF_T() - tq;\n      tc_fence_after(); tq = PROF_T();")
    rep("      const float m_new = fmaxf(", "      pmx += PROF_T() - tq;\n      const float m_new = fmaxf(")
    rep("        mbar_wait(bars + B_PVDONE + hf, (j - 1) & 1);              // every P V into this accumulator has landed",
        "        tq = PROF_T(); ++nrs; mbar_wait(bars + B_PVDONE + hf, (j - 1) & 1);")
    rep("        for (int i = 0; i < 4; ++i) l4[i] *= f;\n      }", "        for (int i = 0; i < 4; ++i) l4[i] *= f;\n        prs += PROF_T() - tq;\n      }")
    rep("      if (j > 0) mbar_wait(bars + B_PVDONE + hf, (j - 1) & 1);",
        "      tq = PROF_T(); if (j > 0) mbar_wait(bars + B_PVDONE + hf, (j - 1) & 1); pwpv += PROF_T() - tq;")
    rep("    const float l = (l4[0] + l4[1]) + (l4[2] + l4[3]);", "    const long long pend1 = PROF_T();\n    const float l = (l4[0] + l4[1]) + (l4[2] + l4[3]);")
    rep("      mbar_wait(bars + B_OFULL, 0);", "      tq = PROF_T(); mbar_wait(bars + B_OFULL, 0); pwo = PROF_T() - tq;")
    # thread 64 = lane 0 of warp 2 (lower-half warpgroup), thread 192 = lane 0 of warp 6 (upper half): rows cta and 512 + cta
    rep("  tc_fence_before();\n  __syncthreads();\n  if (warp == 1) { tc_fence_after(); tmem_dealloc<256>(tmem); }\n}",
        "  tc_fence_before();\n  __syncthreads();\n  if (warp == 1) { tc_fence_after(); tmem_dealloc<256>(tmem); }\n}")
    rep("        }\n      }\n    }\n  }\n  tc_fence_before();",
        "        }\n      }\n    }\n    if (threadIdx.x == 64 || threadIdx.x == 192) {\n      const int cta = blockIdx.y * gridDim.x + blockIdx.x;\n      if (cta < 512) {\n"
        "        long long* o = g_attn_prof[cta + (threadIdx.x == 192 ? 512 : 0)];\n        o[0] = PROF_T() - pstart; o[1] = pend1 - pstart; o[2] = pmx; o[3] = pws; o[4] = pld; o[5] = pwpv; o[6] = prs; o[7] = nrs; o[8] = pwo;\n"
        "        unsigned smid; asm(\"mov.u32 %0, %%smid;\" : \"=r\"(smid)); o[9] = smid; o[10] = pstart;\n      }\n    }\n  }\n  tc_fence_before();")
    rep('extern "C" int nhmc_attn_fwd',
        'extern "C" int nhmc_attn_prof_dump(long long* host) {\n  return (int)cudaMemcpyFromSymbol(host, g_attn_prof, sizeof(g_attn_prof));\n}\n\nextern "C" int nhmc_attn_fwd')
    return s


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pad", type=int, default=0)
    a = ap.parse_args()
    out_dir = os.path.join(ROOT, "tools", "exp", "libvar")
    os.makedirs(out_dir, exist_ok=True)
    src = os.path.join(out_dir, "attn_prof_gen.cu")
    open(src, "w").write(instrument(open(os.path.join(CSRC, "attn.cu")).read(), a.pad))
    objs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".o") and f != "attn.o"]
    obj = os.path.join(out_dir, "attn_prof_gen.o")
    arch = ["-gencode", "arch=compute_100a,code=sm_100a"]
    subprocess.check_call(["nvcc", "-O3", "-std=c++17", "-lineinfo"] + arch + ["-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-I", CSRC,
                           "-I", os.path.join(ROOT, "include"), "-c", src, "-o", obj])
    subprocess.check_call(["nvcc"] + arch + ["-shared", "-o", os.path.join(out_dir, "libnhmc_attnprof.so"), obj] + objs + ["-cudart", "static"])
    print("built", os.path.join(out_dir, "libnhmc_attnprof.so"))


if __name__ == "__main__":
    main()
