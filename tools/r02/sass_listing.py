"""profiles/r02_sass_opcodes.md: per-kernel opcode counts of libmoolib_b200.so (cuobjdump -sass), demangled with c++filt.
   python tools/r02/sass_listing.py > profiles/r02_sass_opcodes.md"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
txt = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "moolib_b200", "lib", "libmoolib_b200.so")],
                     capture_output=True, text=True, check=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)
out = ["# SASS opcode listing of `moolib_b200/lib/libmoolib_b200.so` (round 2)\n",
       "Command: `python tools/r02/sass_listing.py` = `cuobjdump -sass moolib_b200/lib/libmoolib_b200.so` (all cubins `sm_100a`, built by",
       "`python moolib_b200/build.py`), opcodes counted per kernel: async-copy / mbarrier / load-store / fence opcodes listed, the rest summed.",
       "",
       "What to read off: the HP-B kernels drive the TMA engine in linear (`cp.async.bulk`) mode -- `UBLKCP.G.S` (global->shared), `UBLKCP.S.G`",
       "(shared->global), completing on mbarriers (`SYNCS.ARRIVE.TRANS64`, `SYNCS.PHASECHK.TRANS64.TRYWAIT`); there are no tensor maps",
       "(`UTMALDG/UTMASTG` absent: the copies are 1-D spans).  Their LDG side moves 16 bytes per lane with `LDG.E.NA.128.CONSTANT` / `STG.E.NA.128`",
       "(L1 no-allocate).  The HP-A kernels move 16-byte `LD.E.NA.128` / `ST.E.128` over peer mappings and synchronise with system-scope",
       "`ST.E.STRONG.SYS` (release), `LD.E.STRONG.SYS` (relaxed polling) and `MEMBAR.*.SYS` fences.  No kernel contains a tensor-core opcode",
       "(`HMMA`, `UTCMMA`, `UTCHMMA`, `QMMA` ...): neither hot path is a contraction.\n",
       "| kernel | async bulk copy (TMA engine) | mbarrier | loads / stores | fences, cache control | total instructions |",
       "|---|---|---|---|---|---|"]
for f in funcs[1:]:
    name = f.split("\n")[0].strip()
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    dem = re.sub(r"\(anonymous namespace\)::", "", dem)
    dem = re.sub(r"^void ", "", dem)
    dem = re.sub(r"\(.*$", "", dem)
    ops = re.findall(r"^\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)", f, flags=re.M)
    c = collections.Counter(ops)
    pick = lambda pred: ", ".join(f"`{k}`x{v}" for k, v in sorted(c.items()) if pred(k)) or "-"  # noqa: E731
    assert not [k for k in c if "MMA" in k], "tensor-core opcode in " + dem
    out.append("| `%s` | %s | %s | %s | %s | %d |" % (
        dem, pick(lambda k: k.startswith("UBLKCP") or k.startswith("UTMA")), pick(lambda k: k.startswith("SYNCS")),
        pick(lambda k: re.match(r"(LDG|STG|LD|ST)(\.|$)", k) is not None),
        pick(lambda k: k.startswith("MEMBAR") or k.startswith("FENCE") or "ERRBAR" in k or k.startswith("CCTL")), sum(c.values())))
out.append("\ncubin architectures found: %s; kernels: %d" % (sorted(set(re.findall(r"arch = (sm_\w+)", txt))), len(funcs) - 1))
print("\n".join(out))
