#!/usr/bin/env python
"""Per-kernel SASS comparison of two object files / libraries (cuobjdump -sass).

Used when sources change AFTER a binary has been validated on the GPU and no GPU time is left: every kernel that existed in
the validated build must come out instruction-for-instruction identical, so that only genuinely new kernels are unvalidated.

    python scripts/sass_diff.py OLD.o NEW.o          # exit 1 if any kernel present in both differs
"""

import hashlib
import re
import subprocess
import sys


def kernels(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    funcs, name, body = {}, None, []
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if name:
                funcs[name] = hashlib.sha256("\n".join(body).encode()).hexdigest()
            name, body = m.group(1), []
        elif name and re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", line):  # instruction lines only (address + mnemonic), not the encoding-only lines
            body.append(re.sub(r"/\* 0x[0-9a-f]+ \*/", "", line).strip())
    if name:
        funcs[name] = hashlib.sha256("\n".join(body).encode()).hexdigest()
    return funcs


def main(old, new):
    a, b = kernels(old), kernels(new)
    changed = sorted(k for k in a if k in b and a[k] != b[k])
    print(f"{old} -> {new}: {len(a)} kernels before, {len(b)} after; {len(set(a) & set(b)) - len(changed)} identical, "
          f"{len(changed)} changed, {len(set(b) - set(a))} new, {len(set(a) - set(b))} removed")  # fmt: skip
    for k in changed:
        print("  CHANGED", k)
    for k in sorted(set(a) - set(b)):
        print("  REMOVED", k)
    return 1 if changed else 0


if __name__ == "__main__":
    sys.exit(main(*sys.argv[1:3]))
