"""Regenerates microclimf_b200/tables.py from the reference's bundled soil tables
(/root/reference/data/soilparameters.rda, soilparamsp.rda).  Build container only."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from microclimf_b200.rdata import dataframe_columns, read_rda  # noqa: E402

HEADER = '''"""Soil parameter tables of the reference package (datasets `soilparameters` and `soilparamsp`,
documented in /root/reference/R/data.R; values read from data/soilparameters.rda and data/soilparamsp.rda
by tools/make_tables.py).  `.soilinit` (R/internal.R:304-335) looks soil types up in `soilparameters` by
`Number`; `.sortsoilc(method = "P")` and `runpointmodel` index `soilparamsp` by row (R/internal.R:341-357,
R/Cppwrappers.R:118-126)."""
'''


def fmt(d):
    out = []
    for k, v in d.items():
        if isinstance(v, np.ndarray):
            out.append(f'    "{k}": {[float(x) if v.dtype != np.int32 else int(x) for x in v]!r},')
        else:
            out.append(f'    "{k}": {list(v)!r},')
    return "\n".join(out)


sp = dataframe_columns(read_rda("/root/reference/data/soilparameters.rda")["soilparameters"])
spp = dataframe_columns(read_rda("/root/reference/data/soilparamsp.rda")["soilparamsp"])
with open(os.path.join(ROOT, "microclimf_b200", "tables.py"), "w") as f:
    f.write(f"{HEADER}\nSOILPARAMETERS = {{\n{fmt(sp)}\n}}\n\nSOILPARAMSP = {{\n{fmt(spp)}\n}}\n")
