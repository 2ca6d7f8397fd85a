"""Compare the SASS of two builds function by function (instruction text and encodings,
whitespace and addresses ignored):  python scripts/compare_sass.py old.sass new.sass
where the inputs come from `cuobjdump -sass lib.so`.  Functions only present in the new
build are listed as added.  Template instances of fused_stage_ws_kernel are matched modulo
the trailing `AL = false` argument added in round 1."""
import re
import sys


def functions(path):
    out, name, body = {}, None, []
    for line in open(path):
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if name:
                out[name] = body
            name, body = m.group(1), []
        elif name is not None:
            if line.startswith("Fatbin"):
                out[name] = body
                name = None
                continue
            text = re.sub(r"/\*[0-9a-f]{4}\*/", "", line)          # instruction address
            text = " ".join(text.split())
            if text:
                body.append(text)
    if name:
        out[name] = body
    return out


def norm(name):
    return re.sub(r"(fused_stage_ws_kernelI\w+?)Lb0EEEvNS", r"\1EEvNS", name)


if __name__ == "__main__":
    old, new = functions(sys.argv[1]), functions(sys.argv[2])
    new_n = {norm(k): v for k, v in new.items()}
    missing = [k for k in old if k not in new_n]
    differing = [k for k in old if k in new_n and old[k] != new_n[k]]
    added = [k for k in new if norm(k) not in old]
    print(f"{len(old)} functions before, {len(new)} after: {len(missing)} missing, "
          f"{len(differing)} differing, {len(added)} added")
    for k in missing + differing:
        print("  ", k)
    sys.exit(1 if missing or differing else 0)
