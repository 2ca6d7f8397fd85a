"""Poor man's pyflakes (not in this image): report global names that functions load but
their module never defines -- catches typos in code paths that only run on the GPU box."""
import builtins
import dis
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def code_objects(code):
    yield code
    for c in code.co_consts:
        if isinstance(c, types.CodeType):
            yield from code_objects(c)


def check(path):
    src = open(path).read()
    top = compile(src, path, "exec")
    # every name the module itself may bind at top level (no execution needed)
    defined = {"__file__", "__name__", "__doc__"}
    for ins in dis.get_instructions(top):
        if ins.opname in ("STORE_NAME", "STORE_GLOBAL", "IMPORT_NAME", "IMPORT_FROM"):
            defined.add(str(ins.argval).split(".")[0])
    for code in code_objects(top):
        for ins in dis.get_instructions(code):
            if ins.opname == "STORE_GLOBAL":
                defined.add(ins.argval)
    bad = set()
    for code in code_objects(top):
        if code is top:
            continue
        for ins in dis.get_instructions(code):
            if ins.opname in ("LOAD_GLOBAL", "LOAD_NAME"):
                n = ins.argval
                if n not in defined and not hasattr(builtins, n):
                    bad.add((code.co_name, n))
    return sorted(bad)


if __name__ == "__main__":
    status = 0
    for path in sys.argv[1:]:
        for fn, n in check(path):
            print(f"{path}: {fn}: undefined global '{n}'")
            status = 1
    sys.exit(status)
