#!/usr/bin/env python3
"""oracle/gen_ref_headers.py REF OUTDIR SRC... - build-time stand-ins for two headers the reference's build generates.

TEST INFRASTRUCTURE (oracle/Makefile runs it where /root/reference exists; the output goes to oracle/_ref/gen/, which is
git-ignored).  The reference's configure / Gen_fmgrtab.pl / generate-errcodes.pl are not run; instead:

  utils/errcodes.h    one #define per line of the reference's src/backend/utils/errcodes.txt (the SQLSTATE table
                      generate-errcodes.pl reads): `sqlstate E/W/S ERRCODE_NAME ...` -> MAKE_SQLSTATE('x','x','x','x','x')
  utils/fmgrprotos.h  `extern Datum name(PG_FUNCTION_ARGS);` for every fmgr-style function DEFINED in the reference
                      sources that are compiled into oracle/_ref/libexec_ref.so (their definitions need prototypes)
"""
import os
import re
import sys


def main():
    ref, out, srcs = sys.argv[1], sys.argv[2], sys.argv[3:]
    os.makedirs(os.path.join(out, "utils"), exist_ok=True)
    lines = ["/* derived at build time from the reference's src/backend/utils/errcodes.txt */", "#ifndef ERRCODES_H_STANDIN",
             "#define ERRCODES_H_STANDIN"]
    for line in open(os.path.join(ref, "src/backend/utils/errcodes.txt")):
        p = line.split()
        if len(p) >= 3 and len(p[0]) == 5 and p[1] in ("E", "W", "S") and p[2].startswith("ERRCODE_"):
            lines.append("#define %s MAKE_SQLSTATE(%s)" % (p[2], ",".join("'%s'" % c for c in p[0])))
    lines.append("#endif")
    open(os.path.join(out, "utils/errcodes.h"), "w").write("\n".join(lines) + "\n")
    names = set()
    for s in srcs:
        names.update(re.findall(r"^([A-Za-z_0-9]+)\(PG_FUNCTION_ARGS\)", open(s).read(), re.M))
    protos = ["/* derived at build time: fmgr-style functions defined in " + ", ".join(os.path.basename(s) for s in srcs) + " */",
              "#ifndef FMGRPROTOS_H", "#define FMGRPROTOS_H", '#include "fmgr.h"']
    protos += ["extern Datum %s(PG_FUNCTION_ARGS);" % n for n in sorted(names)]
    protos.append("#endif")
    open(os.path.join(out, "utils/fmgrprotos.h"), "w").write("\n".join(protos) + "\n")


if __name__ == "__main__":
    main()
