"""Per-kernel SASS census of libcdmft_b200.so -> profiles/r2_sass_grep.txt (runs without a GPU):
bulk-async copies (UBLKCP = cp.async.bulk), 2-D TMA tensor loads (UTMALDG), mbarrier ops (SYNCS), DFMA, LDS, LDG.
  python tools/sass_census.py"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ("UBLKCP", "UTMALDG", "SYNCS", "DFMA", "LDS", "LDG")


def main():
    lib = os.path.join(ROOT, "cdmft_lanc_ed_b200", "libcdmft_b200.so")
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    cur, cnt = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.search(r"Function : _ZN2cb(\S+)", line)
        if m:
            cur = re.sub(r"^\d+", "", m.group(1))
            cnt[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        for k in KEYS:
            if re.search(r"\b" + k + r"\b", line):
                cnt[cur][k] += 1
    path = os.path.join(ROOT, "profiles", "r2_sass_grep.txt")
    with open(path, "w") as f:
        f.write("# cuobjdump -sass cdmft_lanc_ed_b200/libcdmft_b200.so (tools/sass_census.py): per kernel, count of bulk-async copies "
                "(UBLKCP = cp.async.bulk), 2-D TMA tensor loads (UTMALDG = cp.async.bulk.tensor), mbarrier ops (SYNCS), DFMA, "
                "shared-memory loads (LDS), global loads (LDG)\n")
        for name in sorted(cnt):
            f.write(name + " " + " ".join(f"{k}={cnt[name][k]}" for k in KEYS) + "\n")
    print(f"{len(cnt)} kernels -> {path}")


if __name__ == "__main__":
    main()
