"""Compile and run tools/attn_tr_emu.cpp around the TR = 1 blocks cut out of paged_attn.cu (see the
header of the .cpp).  Exit status 0 = the kernel's own statements, run by 32 host threads with
emulated warp collectives, reproduce softmax(QK^T)V on every case."""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def extract(src: str, name: str, family: str = "tr") -> str:
    m = re.search(r"// \[%s-emu:%s begin\][^\n]*\n(.*?)\n\s*// \[%s-emu:%s end\]" % (family, name, family, name), src, re.S)
    if not m:
        raise RuntimeError(f"marker block {name} not found in paged_attn.cu")
    return m.group(1)


def main() -> int:
    src = open(os.path.join(ROOT, "scalellm_b200", "csrc", "paged_attn.cu")).read()
    rc = 0
    # first the default (GPU-validated) blocks: they validate the emulated warp instructions;
    # then the transposed-tile blocks
    for tr, family in ((0, "def"), (1, "tr")):
        with tempfile.TemporaryDirectory() as tmp:
            for name in ("load_q", "tile", "finalize"):
                text = extract(src, name, family)
                if family == "def":  # the marker's comment continues on two more lines
                    text = re.sub(r"^\s*// GPU-validated, so they validate[^\n]*\n", "", text)
                with open(os.path.join(tmp, f"attn_tr_emu_{name}.inc"), "w") as f:
                    f.write(text + "\n")
            exe = os.path.join(tmp, "attn_tr_emu")
            r = subprocess.run(["g++", "-O1", "-std=c++17", "-pthread", "-Wno-unknown-pragmas", f"-DEMU_TR={tr}",
                                "-I", tmp, os.path.join(ROOT, "tools", "attn_tr_emu.cpp"), "-o", exe],
                               capture_output=True, text=True)
            if r.returncode != 0:
                sys.stderr.write(r.stderr[-4000:])
                return 2
            r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
            sys.stdout.write(f"-- {'transposed tile (TR = 1)' if tr else 'default tile (GPU-validated)'}\n" + r.stdout)
            sys.stderr.write(r.stderr[-2000:])
            rc |= r.returncode
    return rc


if __name__ == "__main__":
    sys.exit(main())
