"""Build alternative copies of libsrs_ctr.so with extra -D flags on ONE kernel source (tuning experiments):

    python profiles/exp/build_variants.py din_rtp.cu nowd:-DRTP_NO_WATCHDOG lazy0:-DRTP_LAZY_NS=0 ...

-> sparrowrecsys_b200/variants/libsrs_ctr_<name>.so, selected at run time with SRS_CTR_LIB=<path>.
The other objects are taken from sparrowrecsys_b200/build/ (run the normal build first)."""
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from sparrowrecsys_b200 import build as B   # noqa: E402


def main():
    src = sys.argv[1]
    B.build()
    out_dir = os.path.join(B.HERE, "variants")
    os.makedirs(out_dir, exist_ok=True)
    nvcc = B.nvcc_path()
    base = os.path.basename(src)[:-3]
    objs = [o for o in glob.glob(os.path.join(B.HERE, "build", "*.o")) if os.path.basename(o) != base + ".o"]
    procs = []
    for spec in sys.argv[2:]:
        name, flags = spec.split(":", 1)
        obj = os.path.join(out_dir, "%s_%s.o" % (base, name))
        cmd = [nvcc, *B.ARCH, "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
               "--extended-lambda", *flags.split(","), "-c", os.path.join(B.CSRC, src), "-o", obj]
        procs.append((name, obj, subprocess.Popen(cmd)))
    for name, obj, p in procs:
        assert p.wait() == 0, name
        lib = os.path.join(out_dir, "libsrs_ctr_%s.so" % name)
        subprocess.check_call([nvcc, *B.ARCH, "-shared", "-o", lib, obj, *objs])
        print(lib)


if __name__ == "__main__":
    main()
