"""Function-by-function SASS comparison of two builds of libmmr_b200.so (no GPU needed).

    python tools/sass_diff.py OLD.so [NEW.so]        # NEW defaults to the in-tree library

Used to show that adding an opt-in kernel variant (a new template parameter with a default, a new loader ...) leaves the
instruction stream of the kernels that already ran on a GPU untouched.  Names are compared after dropping defaulted
trailing template arguments (`...ELb0EEEv` -> `...EEEv`), code after dropping the address column."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sass(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    fns, cur = {}, None
    for ln in out.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1).replace("ELb0EEEv", "EEEv")
            fns[cur] = []
        elif cur is not None:
            t = re.sub(r"/\*.*?\*/", "", ln).strip()
            if t:
                fns[cur].append(t)
    return fns


def main():
    old = sass(sys.argv[1])
    new = sass(sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "multimodalrouting_b200", "csrc", "libmmr_b200.so"))
    same = 0
    for name, code in sorted(old.items()):
        if name not in new:
            print("MISSING  ", name)
        elif new[name] == code:
            same += 1
        else:
            other = new[name]
            n = sum(1 for a, b in zip(code, other) if a != b) + abs(len(code) - len(other))
            perm = sorted(re.sub(r"\bU?R\d+\b", "R", x) for x in code) == sorted(re.sub(r"\bU?R\d+\b", "R", x) for x in other)
            print(f"DIFFERENT {name}: {n} of {len(code)} instructions" + ("  (same instructions up to register names / order)" if perm else ""))
    print(f"{same} of {len(old)} kernels identical; {len(set(new) - set(old))} new kernels")


if __name__ == "__main__":
    main()
