"""Per-kernel counts of the SASS mnemonics that prove the hardware path (tcgen05 = UTC*MMA / UTCBAR / LDTM,
TMA = UTMALDG, TMEM alloc = UTCATOMSWS, mbarrier = SYNCS, clusters / DSMEM = UCGABAR / MAPA / LDS..., vector
global access = LDG/STG .128).  Runs on the CPU: `python scripts/sass_evidence.py > profiles/sass_evidence.txt`."""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "blades_b200", "_cuda.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", so], capture_output=True, text=True).stdout
usage = dict(re.findall(r"Function (\S+):\n\s*(REG:\d+ .*)", res))
KEY = re.compile(r"^(UTC\w*MMA|UTCBAR|UTCATOMSWS|UTMALDG|UTMAPF|UTMACCTL|LDTM|STTM|SYNCS|UCGABAR|MAPA|CCTL|REDG|RED|"
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
