"""Static SASS instruction count per CUDA source line for one kernel (nvdisasm -g -c output)."""
import collections
import re
import sys


def main(path, kernel_substr, top=40):
    lines = open(path).read().split("\n")
    in_k = False
    cur = ("?", 0)
    counts = collections.Counter()
    total = 0
    for ln in lines:
        if ln.startswith(".text.") or re.match(r"^\s*\.section\s+\.text\.", ln):
            in_k = kernel_substr in ln
            continue
        if not in_k:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        if re.match(r"^\s+/\*[0-9a-f]{4}\*/\s+\S", ln):
            counts[cur] += 1
            total += 1
    print("static SASS instructions:", total)
    src_cache = {}
    for (f, l), n in counts.most_common(top):
        text = ""
        for cand in ("marl_range_flocking_b200/csrc/" + f,):
            try:
                if cand not in src_cache:
                    src_cache[cand] = open(cand).read().split("\n")
                text = src_cache[cand][l - 1].strip()[:90]
            except Exception:
                pass
        print(f"{n:5d}  {f}:{l:<4d} {text}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40)
