"""Parse include/deepatlas_b200.h into {name: (arg codes, return kind)} for ABI cross-checks."""
import os
import re

HEADER = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "deepatlas_b200.h")


def parse_header(path=HEADER):
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"(int64_t|int|const char\*)\s+(da_\w+)\s*\(([^)]*)\)\s*;", src):
        ret, name, args = m.groups()
        codes = ""
        for a in [a.strip() for a in args.split(",")]:
            if a in ("void", ""):
                continue
            if a.startswith("da_stream_t"):
                codes += "s"
            elif "*" in a:
                codes += "p"
            elif a.startswith("int64_t"):
                codes += "l"
            elif a.startswith("int"):
                codes += "i"
            elif a.startswith("float"):
                codes += "f"
            elif a.startswith("double"):
                codes += "d"
            else:
                raise ValueError(f"{name}: cannot classify argument '{a}'")
        out[name] = (codes, ret)
    return out
