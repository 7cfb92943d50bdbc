"""Extracts the XSPEC model table of the reference (src/modelfiles/lmodel_relxill_public.dat and
lmodel_relxill_devel.dat) into tests/golden/lmodel_layout.json: per model the C symbol XSPEC resolves (column 5
without the `c_` prefix), the model type (add / con) and the parameters in file order with their default, hard
limits and soft limits ($switch entries included, as the wrapper passes them).  Run in the build container:
    python tests/golden/make_lmodel_layout.py
The fixture pins the drop-in boundary (parameter order, defaults, symbol names) on machines without the reference."""
import json
import os

SRC = "/root/reference/src/modelfiles"
HERE = os.path.dirname(os.path.abspath(__file__))


def parse(path):
    models, cur, left = {}, None, 0
    for ln, line in enumerate(open(path), 1):
        t = line.split()
        if not t:
            continue
        if left == 0:
            name, npar, sym, mtype = t[0], int(t[1]), t[4], t[5]
            cur = dict(symbol=sym[2:] if sym.startswith("c_") else sym, type=mtype, file=os.path.basename(path), line=ln, params=[])
            models[name] = cur
            left = npar
        else:
            if t[0].startswith("$"):
                cur["params"].append(dict(name=t[0][1:], default=float(t[1]), switch=True))
            else:
                vals = [float(x) for x in t[2:8]]
                cur["params"].append(dict(name=t[0], unit=t[1].strip('"'), default=vals[0], hard_min=vals[1], soft_min=vals[2],
                                          soft_max=vals[3], hard_max=vals[4], switch=False))
            left -= 1
    return models


def main():
    out = {}
    for f in ("lmodel_relxill_public.dat", "lmodel_relxill_devel.dat"):
        out.update(parse(os.path.join(SRC, f)))
    json.dump(out, open(os.path.join(HERE, "lmodel_layout.json"), "w"), indent=1, sort_keys=True)
    for k, v in out.items():
        print(k, v["symbol"], v["type"], len(v["params"]))


if __name__ == "__main__":
    main()
