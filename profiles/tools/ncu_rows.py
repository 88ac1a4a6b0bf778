"""ncu --csv --metrics log -> one line per launch: id, kernel, metric=value ..."""
import csv, sys
from collections import OrderedDict
rows = [r for r in csv.reader(open(sys.argv[1])) if r and r[0] != "" and not r[0].startswith("==")]
hdr = next(r for r in rows if "Kernel Name" in r)
iK, iM, iV, iID = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
d = OrderedDict()
for r in rows:
    if r is hdr or len(r) <= iV or r[iID] == "ID":
        continue
    d.setdefault((r[iID], r[iK][:30]), {})[r[iM]] = r[iV]
for (i, k), m in d.items():
    print(i, k, " ".join(f"{a.split('__')[-1][:30]}={b}" for a, b in m.items()))
