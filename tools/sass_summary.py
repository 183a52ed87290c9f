"""SASS evidence for profiles/: instruction mix of one kernel of libspring_b200.so (cuobjdump -sass) and the lines that
show how it touches memory (vector width, cache operators / hints, atomics, warp collectives, TMA).
usage: sass_summary.py <kernel-substring> [<kernel-substring> ...]"""
import collections, re, subprocess, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "spring_b200", "libspring_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
funcs, cur = {}, None
for ln in txt.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1); funcs[cur] = []; continue
    if cur and re.search(r"/\*[0-9a-f]{4,}\*/", ln):
        ins = re.sub(r"/\*[0-9a-f]+\*/", "", ln).strip().rstrip(";").strip()
        if ins:
            funcs[cur].append(ins)
for want in sys.argv[1:]:
    for name, body in funcs.items():
        if want not in name:
            continue
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        print(f"== {dem[:160]}")
        ops = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", i).split()[0] for i in body)
        print(f"   {len(body)} SASS instructions")
        groups = {"global loads": r"^LDG", "global stores": r"^STG", "atomics / reductions to memory": r"^(ATOMG|ATOM|RED)\b|^ATOMG|^REDG",
                  "shared loads / stores": r"^(LDS|STS)", "local (spill) loads / stores": r"^(LDL|STL)", "warp collectives": r"^(SHFL|VOTE|REDUX|MATCH|WARPSYNC)",
                  "population count / bit reverse / funnel shift": r"^(POPC|BREV|SHF|FLO)", "TMA / bulk copy / mbarrier": r"^(UBLKCP|UTMA|SYNCS)",
                  "cache control (prefetch)": r"^CCTL", "constant-bank loads": r"^(LDC|ULDC)"}
        for g, rx in groups.items():
            sel = {k: v for k, v in ops.items() if re.search(rx, k)}
            if sel:
                print(f"   {g:48s} {sum(sel.values()):5d}   " + ", ".join(f"{k} x{v}" for k, v in sorted(sel.items(), key=lambda kv: -kv[1])[:8]))
        seen = set()
        print("   memory / collective instructions as emitted (first of each kind):")
        for i in body:
            op = re.sub(r"^@!?U?P\d+\s+", "", i).split()[0]
            if re.search(r"^(LDG|STG|ATOM|RED|REDUX|MATCH|UBLKCP|UTMA|SYNCS|CCTL|LDL|STL)", op) and op not in seen:
                seen.add(op); print("      " + i[:150])
