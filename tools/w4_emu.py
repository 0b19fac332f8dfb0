"""Compile and run tools/w4_emu.cpp around the blocks cut out of w4a16.cu / common.cuh (see the
header of the .cpp), once per kernel variant.  Exit status 0 = the kernel's own role code, run by
one host thread per warp over emulated mbarriers / copies / tensor memory / tensor pipe, keeps every
protocol rule on every CTA share tried, for the default kernel and every B200_W4_VARIANT."""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VARIANTS = (0, 1, 2, 4, 6, 10, 14, 16, 20, 24, 28)
# (variant, rows-of-the-batch tile): the default kernel also at 16 and 128 rows (128: 3-stage
# activation ring, 4 TMEM slots), two variants at 16 rows
CONFIGS = tuple((v, 64) for v in VARIANTS) + ((0, 16), (0, 128), (2, 16), (16, 16))


def extract(src: str, name: str) -> str:
    m = re.search(r"// \[w4-emu:%s begin\][^\n]*\n(.*?)\n[^\n]*// \[w4-emu:%s end\]" % (name, name), src, re.S)
    if not m:
        raise RuntimeError(f"marker block {name} not found")
    return m.group(1)


def main(rounds: int = 12) -> int:
    csrc = os.path.join(ROOT, "scalellm_b200", "csrc")
    w4 = open(os.path.join(csrc, "w4a16.cu")).read()
    common = open(os.path.join(csrc, "common.cuh")).read()
    rc = 0
    with tempfile.TemporaryDirectory() as tmp:
        blocks = {"plan": extract(common, "plan"), "cfg": extract(w4, "cfg"), "init": extract(w4, "init"),
                  "roles": extract(w4, "roles")}
        # the cfg block's first comment line continues onto a second line in the source
        blocks["cfg"] = re.sub(r"^// role code of the kernel for the host\)\n", "", blocks["cfg"])
        for name, text in blocks.items():
            with open(os.path.join(tmp, f"w4_emu_{name}.inc"), "w") as f:
                f.write(text + "\n")
        for var, mt in CONFIGS:
            exe = os.path.join(tmp, f"w4_emu_{var}_{mt}")
            r = subprocess.run(["g++", "-O1", "-std=c++17", "-pthread", "-Wno-unknown-pragmas", f"-DEMU_VAR={var}",
                                f"-DEMU_MT={mt}", "-I", tmp, os.path.join(ROOT, "tools", "w4_emu.cpp"), "-o", exe],
                               capture_output=True, text=True)
            if r.returncode != 0:
                sys.stderr.write(r.stderr[-6000:])
                return 2
            r = subprocess.run([exe, str(rounds)], capture_output=True, text=True, timeout=900)
            sys.stdout.write(r.stdout)
            sys.stderr.write(r.stderr[-2000:])
            rc |= r.returncode
    return rc


if __name__ == "__main__":
    sys.exit(main(int(sys.argv[1]) if len(sys.argv) > 1 else 12))
