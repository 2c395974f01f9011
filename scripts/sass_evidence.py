"""Per-kernel counts of the SASS mnemonics that prove the hardware path (tcgen05 = UTC*MMA / UTCBAR / LDTM,
TMA = UTMALDG, TMEM alloc = UTCATOMSWS, mbarrier = SYNCS, clusters / DSMEM = UCGABAR / MAPA / LDS..., vector
global access = LDG/STG .128, NVLS multimem.ld_reduce = LDGMC, multimem.st = STG.E[.128].STRONG.SYS on the multicast
address, bulk copies = UBLKCP).  ``--listings DIR`` also writes the full gzip'ed SASS of the tcgen05 / TMA / multimem /
headline kernels.  Runs on the CPU: `python scripts/sass_evidence.py > profiles/sass_evidence.txt`."""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "blades_b200", "_cuda.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", so], capture_output=True, text=True).stdout
usage = dict(re.findall(r"Function (\S+):\n\s*(REG:\d+ .*)", res))
KEY = re.compile(r"^(LDGMC|STG\.E(\.128)?\.STRONG\.SYS|IMAD$|IMAD\.WIDE|UBLKCP|UTC\w*MMA|UTCBAR|UTCATOMSWS|UTMALDG|UTMAPF|UTMACCTL|LDTM|STTM|SYNCS|UCGABAR|MAPA|CCTL|REDG|RED|"
                 r"LDG\.E\.(128|EF|CONSTANT)|LDG\.E\.\w*\.?128|STG\.E\.128|STG\.E\.EF\.128|FMNMX|FFMA|HMMA|ACQBULK|"
                 r"FENCE|MEMBAR|LDS\.128|STS\.128|ST\.E\.128|LD\.E\.128|MUFU)")
cur, counts = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if cur and m:
        op = m.group(1)
        if KEY.match(op):
            counts[cur][op] += 1
for fn, c in counts.items():
    name = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip().replace("(anonymous namespace)::", "").split("(")[0]
    print(f"== {name}    [{usage.get(fn, '').split(' CONSTANT')[0]}]")
    merged = collections.Counter()
    for op, n in c.items():
        merged[re.sub(r"\.(FTZ|RN|STRONG|GPU|SYS|TRANS64|A1T0|ALIGN|NOINC)", "", op)] += n
    for op, n in sorted(merged.items(), key=lambda kv: (-kv[1], kv[0])):
        print(f"   {n:6d}  {op}")

import sys
if "--listings" in sys.argv:
    import gzip
    out_dir = sys.argv[sys.argv.index("--listings") + 1]
    os.makedirs(out_dir, exist_ok=True)
    WANT = ["gram_tcgen05_kernel", "wgrad_tcgen05_kernel", "conv_tcgen05_kernel", "nvls_allreduce_kernel",
            "row_combine_kernelILb1ELb0", "row_combine_kernelILb1ELb1", "coord_select_part_kernelILi80ELi5",
            "coord_select_part_stage_kernelILi80ELi5", "client_bn_nhwc_fwd_cl_kernelILi8", "client_bn_nhwc_bwd_cl_kernelILi8",
            "gram_krum_kernel", "gram_iter_kernel"]
    blocks, cur = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            blocks[cur] = []
        if cur:
            blocks[cur].append(line)
    for w in WANT:
        for fn, lines in blocks.items():
            if w in fn:
                with gzip.open(os.path.join(out_dir, w + ".sass.gz"), "wt") as f:
                    f.write("\n".join(lines) + "\n")
                break
