#!/usr/bin/env python3
"""Build the UNMODIFIED reference `spacedust` binary into oracle/_ref/ (test infrastructure).

This is the recipe the task calls "oracle/_ref": the reference's own C++ sources are compiled
*where they lie* under /root/reference with plain g++ driven by a Makefile this script writes.
The reference's CMake build system is NOT run. The only "generated code" the reference needs is
the byte-embedding of its data/*.sh, *.out and *.lib resources into C arrays (what
lib/mmseqs/cmake/MMseqsResourceCompiler.cmake does with xxd); this script re-creates that
embedding itself (function `embed`).  No reference source is copied into the repository: the
outputs (generated headers, objects, the binary) live only under oracle/_ref/, which is
git-ignored but travels to the GPU box with the snapshot.

Compile flags mirror the reference's default AVX2 release build
(lib/mmseqs/CMakeLists.txt: -O3 -DNDEBUG -fsigned-char -mavx2 -mfma -mcx16 -std=c++1y -fopenmp).

Usage:  python oracle/build_ref.py [-j N] [--ref /root/reference]
Result: oracle/_ref/spacedust  (≈15 MB), oracle/_ref/BUILD_INFO.json
"""
import argparse
import glob
import json
import os
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))


def embed(src, dst):
    """Same contract as xxd -i + sed 's!unsigned char!static const unsigned char!'."""
    name = os.path.basename(src)
    var = name.replace("/", "_").replace(".", "_").replace("-", "_")
    data = open(src, "rb").read()
    out = ["static const unsigned char %s[] = {" % var]
    for i in range(0, len(data), 16):
        out.append("\t" + ",".join("0x%02x" % b for b in data[i:i + 16]) + ",")
    out.append("};")
    out.append("unsigned int %s_len = %d;" % (var, len(data)))
    with open(dst, "w") as f:
        f.write("\n".join(out) + "\n")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("-j", type=int, default=os.cpu_count() or 4)
    ap.add_argument("--keep-objects", action="store_true")
    ap.add_argument("--force", action="store_true")
    args = ap.parse_args()
    ref = os.path.abspath(args.ref)
    out = os.path.join(HERE, "_ref")
    binary = os.path.join(out, "spacedust")
    if os.path.exists(binary) and not args.force:
        print("oracle/_ref/spacedust already built (use --force to rebuild)")
        return 0
    if not os.path.isdir(os.path.join(ref, "lib", "mmseqs", "src")):
        print("reference tree not found at %s; nothing to build" % ref)
        return 2
    gen = os.path.join(out, "gen")
    obj = os.path.join(out, "obj")
    os.makedirs(gen, exist_ok=True)
    os.makedirs(obj, exist_ok=True)

    mm = os.path.join(ref, "lib", "mmseqs")
    # ---- resource embedding (lib/mmseqs/data/CMakeLists.txt, data/CMakeLists.txt) ----
    resources = []
    resources += [os.path.join(mm, "data", m) for m in
                  ("VTML80.out", "VTML40.out", "nucleotide.out", "blosum62.out", "PAM30.out")]
    resources += sorted(glob.glob(os.path.join(mm, "data", "workflow", "*.sh")))
    for r in sorted(os.listdir(os.path.join(mm, "data", "resources"))):
        if r != "CMakeLists.txt":
            resources.append(os.path.join(mm, "data", "resources", r))
    resources += sorted(glob.glob(os.path.join(ref, "data", "*.sh")))
    for r in resources:
        embed(r, os.path.join(gen, os.path.basename(r) + ".h"))

    # ---- source lists (mirrors the object list of the reference's default build) ----
    cxx = []
    for d in ("alignment", "clustering", "commons", "linclust", "prefiltering", "taxonomy",
              "util", "workflow"):
        cxx += sorted(glob.glob(os.path.join(mm, "src", d, "*.cpp")))
    cxx += [os.path.join(mm, "src", "multihit", f) for f in
            ("besthitperset.cpp", "combinepvalperset.cpp", "Aggregation.cpp", "MultiHitDb.cpp",
             "MultiHitSearch.cpp")]
    cxx.append(os.path.join(mm, "src", "MMseqsBase.cpp"))           # FRAMEWORK_ONLY: no mmseqs.cpp
    cxx += sorted(glob.glob(os.path.join(mm, "lib", "alp", "*.cpp")))
    cxx.append(os.path.join(mm, "lib", "ksw2", "ksw2_extz2_sse.cpp"))
    cxx.append(os.path.join(mm, "lib", "tantan", "tantan.cpp"))
    cxx.append(os.path.join(ref, "src", "spacedust.cpp"))
    cxx.append(os.path.join(ref, "src", "commons", "LocalParameters.cpp"))
    cxx += sorted(glob.glob(os.path.join(ref, "src", "util", "*.cpp")))
    cxx += sorted(glob.glob(os.path.join(ref, "src", "workflow", "*.cpp")))
    cxx.append(os.path.join(ref, "src", "version", "Version.cpp"))
    cc = [os.path.join(mm, "lib", "tinyexpr", "tinyexpr.c"),
          os.path.join(mm, "lib", "microtar", "microtar.c")]
    for d in ("common", "compress", "decompress", "dictBuilder", "deprecated"):
        cc += sorted(glob.glob(os.path.join(mm, "lib", "zstd", "lib", d, "*.c")))

    inc = [os.path.join(ref, "lib"), os.path.join(ref, "src", "commons")]
    inc += [os.path.join(mm, "lib", d) for d in
            ("zstd/lib", "zstd/lib/common", "tinyexpr", "microtar", "tantan", "simde", "", "simd",
             "alp", "ksw2", "xxhash", "ips4o")]
    inc += [gen]
    inc += [os.path.join(mm, "src", d) for d in
            ("alignment", "clustering", "commons", "multihit", "prefiltering", "linclust",
             "taxonomy", "util", "")]
    incflags = " ".join("-I" + i.rstrip("/") for i in inc)
    base = "-O3 -DNDEBUG -fsigned-char -mavx2 -mfma -mcx16 -w"
    defs = ("-DENABLE_IPS4O=1 -DHAVE_POSIX_FADVISE=1 -DHAVE_POSIX_MADVISE=1 -DHAVE_ZLIB=1 "
            "-DOPENMP=1 -DGIT_SHA1=oracle_ref")
    cxxflags = "%s -std=c++1y -fno-exceptions -fopenmp %s %s" % (base, defs, incflags)
    libcxxflags = "%s -std=c++1y %s" % (base, incflags)    # lib/{alp,ksw2,tantan}: exceptions on
    cflags = "%s %s" % (base, incflags)

    # The reference links its own objects first and the MMseqs2 framework as a static archive, so
    # the Spacedust versions of Aggregation/gff2db shadow the framework's same-named members.
    objs = []
    own_objs = []
    fw_objs = []
    rules = []
    for i, s in enumerate(cxx + cc):
        o = os.path.join(obj, "%03d_%s.o" % (i, os.path.basename(s).replace(".", "_")))
        objs.append(o)
        (own_objs if s.startswith(os.path.join(ref, "src") + os.sep) else fw_objs).append(o)
        if s.endswith(".c"):
            rules.append("%s: %s\n\t@/usr/bin/gcc %s -c $< -o $@\n" % (o, s, cflags))
        elif os.sep + os.path.join("mmseqs", "lib") + os.sep in s:
            rules.append("%s: %s\n\t@/usr/bin/g++ %s -c $< -o $@\n" % (o, s, libcxxflags))
        else:
            rules.append("%s: %s\n\t@/usr/bin/g++ %s -c $< -o $@\n" % (o, s, cxxflags))
    mk = os.path.join(out, "Makefile")
    with open(mk, "w") as f:
        f.write("# written by oracle/build_ref.py -- do not edit\n")
        f.write("OWN = %s\n\n" % " \\\n  ".join(own_objs))
        f.write("FW = %s\n\n" % " \\\n  ".join(fw_objs))
        fw = os.path.join(obj, "libfw.a")
        f.write("%s: $(OWN) $(FW)\n\t@rm -f %s\n\t@ar qc %s $(FW)\n"
                "\t@/usr/bin/g++ -O3 -fopenmp -o $@ $(OWN) %s -latomic -lz -lpthread\n\n"
                % (binary, fw, fw, fw))
        f.write("\n".join(rules))
    t0 = time.time()
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)   # the image's /opt/gcc wrapper lacks libgomp.spec
    r = subprocess.run(["make", "-f", mk, "-j", str(args.j), binary], env=env)
    if r.returncode != 0:
        print("reference build FAILED")
        return r.returncode
    if not args.keep_objects:
        for o in objs + [os.path.join(obj, "libfw.a")]:
            try:
                os.remove(o)
            except OSError:
                pass
    info = {"ref": ref, "seconds": round(time.time() - t0, 1), "n_sources": len(cxx) + len(cc),
            "flags": base, "built": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime())}
    with open(os.path.join(out, "BUILD_INFO.json"), "w") as f:
        json.dump(info, f, indent=1)
    print("built %s in %.0f s" % (binary, time.time() - t0))
    return 0


if __name__ == "__main__":
    sys.exit(main())
