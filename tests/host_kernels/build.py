"""Builds tests/host_kernels' emulation library: the factor-family kernel SOURCES (text regions of csrc/*.cu / *.cuh)
compiled as host C++ behind hk_shim.h.  TEST INFRASTRUCTURE for the CPU suite -- nothing under rome.jl_b200/ knows about
it, and it is never used when a GPU test runs."""
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "rome.jl_b200", "csrc")
BANNER = "// ==="
DASHES = "// ---"


def _read(name):
    return open(os.path.join(CSRC, name)).read()


def _between(src, start_marker, end_marker, back_to_banner=True):
    a = src.index(start_marker)
    if back_to_banner:  # include the comment banner the marker sits in
        a = src.rfind("\n", 0, a) + 1
        while True:  # walk back over the comment lines above
            prev = src.rfind("\n", 0, a - 1) + 1
            if src[prev:a].lstrip().startswith("//"):
                a = prev
            else:
                break
    b = src.index(end_marker, a) if end_marker else len(src)
    if end_marker and back_to_banner:
        while True:
            prev = src.rfind("\n", 0, b - 1) + 1
            if src[prev:b].lstrip().startswith("//"):
                b = prev
            else:
                break
    return src[a:b]


def family_region(name):
    s = _read(name)
    a = s.index("namespace rome {") + len("namespace rome {")
    b = s.index("\nint launch_", a)
    return f"\n// ===== {name} =====\n" + s[a:b] + "\n"


def source_text():
    du, ep, se, fk = _read("device_utils.cuh"), _read("eval_pipeline.cuh"), _read("se3_common.cuh"), _read("factor_kernels.cu")
    parts = ['#include "hk_shim.h"\n#include <vector>\n', f'#include "{os.path.join(CSRC, "tables.h")}"\n', "namespace rome {\n"]
    parts.append(_between(du, "constexpr double kPi", "// mbarrier + TMA"))
    parts.append(_between(du, "// Float64 angle helpers", "}  // namespace rome", back_to_banner=True))
    parts.append(_between(ep, "// SE(2) statistics accumulator", "// stage layout (shared by host planning and the kernel)"))
    # the pipeline itself: stage layouts, eval_kernel (producer warp + consumer warps), eval_kernel_w (per-warp rings)
    pipe = ep[ep.index("struct StageLayout {"):ep.index("template <class K>\nint launch_kernel_cfg")]
    pipe = pipe.replace("extern __shared__ __align__(128) unsigned char smem[];",
                        "unsigned char* const smem = hk::g_smem;  /* one block at a time: its dynamic shared memory */")
    pipe, n = re.subn(r'asm volatile\("griddepcontrol\.[a-z_]+;" ::: "memory"\);', "/* griddepcontrol: scheduling hint, no effect on results */;", pipe)
    assert n == 6 and pipe.count("unsigned char* const smem = hk::g_smem") == 2
    parts.append(pipe)
    parts.append(open(os.path.join(HERE, "hk_launch.inc")).read())   # launch_kernel_cfg: run the grid under the emulator
    parts.append(ep[ep.index("template <class Fam, uint32_t kStatic, bool kSample, int FT, bool kRouted = false>\nint launch_ft("):ep.rindex("}  // namespace rome")])
    parts.append(fk[fk.index("struct FamDims {"):fk.index("int launch_eval(")])
    parts.append(_between(fk, "// layout conversion: reference layout", "__global__ void adopt_kernel"))
    parts.append(family_region("fam_pose2.cu"))
    parts.append(family_region("fam_bearingrange.cu"))
    parts.append(family_region("fam_point2.cu"))
    parts.append(family_region("fam_point3.cu"))
    se_body = se[se.index("namespace rome {") + len("namespace rome {"):se.rindex("}  // namespace rome")]
    parts.append("\n// ===== se3_common.cuh =====\n" + se_body)
    parts.append(family_region("fam_se3.cu"))
    parts.append(family_region("fam_se3_partial.cu"))
    parts.append(family_region("fam_se3_ternary.cu"))
    pk = _read("product_kernels.cu")
    parts.append("\n// ===== product_kernels.cu =====\n" + pk[pk.index("constexpr int kProdWarps"):pk.index("\nint launch_product")] + "\n")
    parts.append(open(os.path.join(HERE, "hk_main.inc")).read())
    text = "".join(parts)
    # the few PTX statements of the compiled regions are MUFU approximations: substitute the exact functions
    exact = {"rsqrt": "1.0f / std::sqrt", "lg2": "std::log2", "ex2": "std::exp2"}
    text, n = re.subn(r'asm\("(\w+)\.approx\.ftz\.f32 %0, %1;"\s*:\s*"=f"\((\w+)\)\s*:\s*"f"\((.+?)\)\);',
                      lambda m: f"{m.group(2)} = {exact[m.group(1)]}({m.group(3)});", text)
    assert n >= 2 and "asm(" not in text and "asm volatile" not in text
    return text


def build(outdir):
    os.makedirs(outdir, exist_ok=True)
    src = os.path.join(outdir, "host_kernels.cpp")
    so = os.path.join(outdir, "libhost_kernels.so")
    open(src, "w").write(source_text())
    cmd = ["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-w", "-I", HERE, src, "-o", so, "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("host build of the kernel sources failed:\n" + r.stderr[-6000:])
    return so


def build_peer(outdir):
    """the GPU-side rank barrier (csrc/peer_kernels.cu) as its own host library: its blocks are single warps, so
    __syncthreads is the warp rendezvous and threadIdx the lane; release / acquire accesses become volatile accesses"""
    os.makedirs(outdir, exist_ok=True)
    pk = _read("peer_kernels.cu")
    body = pk[pk.index("struct PeerSlots {"):pk.index("template <class K, class... A>\nstatic int launch_one_warp_pdl")]
    body, n1 = re.subn(r'asm volatile\("st\.release\.sys\.global\.u32 \[%0\], %1;" ::"l"\((.+?)\), "r"\((\w+)\) : "memory"\);',
                       r"*(volatile uint32_t*)(\1) = \2; ++hk::g_progress;", body)
    body, n2 = re.subn(r'asm volatile\("ld\.acquire\.sys\.global\.u32 %0, \[%1\];" : "=r"\((\w+)\) : "l"\((.+?)\) : "memory"\);',
                       r"\1 = *(volatile const uint32_t*)(\2);", body)
    body, n3 = re.subn(r'asm volatile\("griddepcontrol\.[a-z_]+;" ::: "memory"\);', "", body)
    assert n1 == 2 and n2 == 2 and n3 == 6 and "asm" not in body   # signal, wait and the merged barrier kernel
    text = ('#include "hk_shim.h"\n#undef threadIdx\n#define threadIdx (dim3_{(unsigned)(hk::g_cur & 31), 0, 0})\n'
            "#define __syncthreads() hk::collective(0u)\n#define __nanosleep(ns) hk::yield_()\n"
            "namespace rome {\n" + body + open(os.path.join(HERE, "hk_peer_main.inc")).read() + "}\n")
    src, so = os.path.join(outdir, "host_peer.cpp"), os.path.join(outdir, "libhost_peer.so")
    open(src, "w").write(text)
    r = subprocess.run(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-w", "-I", HERE, src, "-o", so], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("host build of peer_kernels.cu failed:\n" + r.stderr[-4000:])
    return so


if __name__ == "__main__":
    import sys
    print(build(sys.argv[1] if len(sys.argv) > 1 else "/tmp/host_kernels"))
