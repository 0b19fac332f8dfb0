"""Compile and run tools/attn_emu.cpp: the whole paged-attention stream kernel, cut out of
paged_attn.cu, executed on the host (see the header of the .cpp) — first the default instantiation
(GPU-validated: it validates the harness), then the opt-in ones (OCC, TR, both)."""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# (OCC, TR, head_dim, simt): simt = 1 runs the CUDA-core kernel (4 warps per CTA) instead of the stream
# kernel: head_dim 128 / 64 are GPU-validated there too (harness check), 32 / 96 are the new ones
CONFIGS = ((0, 0, 128, 0), (1, 0, 128, 0), (0, 1, 128, 0), (1, 1, 128, 0), (0, 0, 64, 0), (1, 1, 64, 0),
           (0, 0, 256, 0), (0, 0, 128, 1), (0, 0, 64, 1), (0, 0, 96, 1), (0, 0, 32, 1))
QUICK = ((0, 0, 128, 0), (1, 1, 128, 0), (1, 1, 64, 0), (0, 0, 128, 1), (0, 0, 96, 1),
         (0, 0, 32, 1))  # the CPU test's subset


def extract(src: str, name: str) -> str:
    m = re.search(r"// \[attn-emu:%s begin\][^\n]*\n(.*?)\n// \[attn-emu:%s end\]" % (name, name), src, re.S)
    if not m:
        raise RuntimeError(f"marker block {name} not found in paged_attn.cu")
    return m.group(1)


def main(quick: bool = False) -> int:
    src = open(os.path.join(ROOT, "scalellm_b200", "csrc", "paged_attn.cu")).read()
    params = re.sub(r"^// kernel for the host\)\n", "", extract(src, "params"))
    kernel = extract(src, "persist")
    combine = extract(src, "combine")
    simt = extract(src, "simt")
    # the two inline-asm statements of the kernel become calls into the harness
    kernel, n1 = re.subn(r'asm volatile\("cp\.async\.ca\.shared\.global \[%0\], \[%1\], 4;" ::"r"\((.*?)\), "l"\((.*?)\)\s*:\s*"memory"\);',
                         r"emu_cp_async4(\1, \2);", kernel, flags=re.S)
    kernel, n2 = re.subn(r'asm volatile\("cp\.async\.wait_all;" ::: "memory"\);', "emu_cp_async_wait_all();", kernel)
    if n1 != 1 or n2 != 1 or "asm" in kernel:
        raise RuntimeError("unexpected inline asm in the kernel block")
    lib_dir = os.path.join(ROOT, "scalellm_b200")
    rc = 0
    with tempfile.TemporaryDirectory() as tmp:
        open(os.path.join(tmp, "attn_emu_params.inc"), "w").write(params + "\n")
        open(os.path.join(tmp, "attn_emu_persist.inc"), "w").write(kernel + "\n")
        open(os.path.join(tmp, "attn_emu_combine.inc"), "w").write(combine + "\n")
        open(os.path.join(tmp, "attn_emu_simt.inc"), "w").write(simt + "\n")
        for occ, tr, hd, simt_k in (QUICK if quick else CONFIGS):
            exe = os.path.join(tmp, f"attn_emu_{occ}{tr}_{hd}_{simt_k}")
            r = subprocess.run(["g++", "-O1", "-std=c++17", "-pthread", "-Wno-unknown-pragmas", f"-DEMU_OCC={occ}",
                                f"-DEMU_TR={tr}", f"-DEMU_D={hd}", f"-DEMU_SIMT={simt_k}", "-I", tmp, os.path.join(ROOT, "tools", "attn_emu.cpp"), "-o", exe,
                                "-L" + lib_dir, "-lb200decode", "-Wl,-rpath," + lib_dir],
                               capture_output=True, text=True)
            if r.returncode != 0:
                sys.stderr.write(r.stderr[-6000:])
                return 2
            env = dict(os.environ, B200_ATTN_OCC=str(occ), B200_ATTN_TR=str(tr))
            if simt_k:
                env["B200_ATTN_IMPL"] = "simt"
            r = subprocess.run([exe], capture_output=True, text=True, timeout=1500, env=env)
            sys.stdout.write(f"-- {'CUDA-core kernel' if simt_k else f'stream kernel OCC={occ} TR={tr}'} head_dim={hd}\n" + r.stdout)
            sys.stderr.write(r.stderr[-2000:])
            rc |= r.returncode
    return rc


if __name__ == "__main__":
    sys.exit(main(quick=len(sys.argv) > 1 and sys.argv[1] == "quick"))
