"""Build helper: turns ew_prelude.cuh (+ the hwy_math.cuh it includes) into a C string literal so libaxb.so can hand the
text to NVRTC. Usage: embed_prelude.py <csrc dir> <out .inc>"""
import sys, os

src_dir, out = sys.argv[1], sys.argv[2]
text = open(os.path.join(src_dir, "ew_prelude.cuh")).read()
inc = open(os.path.join(src_dir, "hwy_math.cuh")).read().replace("#pragma once\n", "")
assert '#include "hwy_math.cuh"' in text
text = text.replace('#include "hwy_math.cuh"', inc).replace("#pragma once\n", "")
with open(out, "w") as f:
    for line in text.split("\n"):
        f.write('"' + line.replace("\\", "\\\\").replace('"', '\\"') + '\\n"\n')
