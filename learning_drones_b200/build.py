"""Build libldrones.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m learning_drones_b200.build [--force] [--verbose]
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libldrones.so")
# the same library with the episode kernel's own bounds checks compiled in (-DLD_CHECKED); tests/test_gpu_soak.py runs it
LIB_CHECKED = os.path.join(HERE, "libldrones_checked.so")
SOURCES = [os.path.join(CSRC, "ldrones.cu")]
DEPS = SOURCES + [os.path.join(CSRC, h) for h in ("ld_step.cuh", "ld_math.cuh", "ld_f32.cuh")] + [os.path.join(ROOT, "include", "ldrones.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    # python-float expressions of the reference are unfused; fused products are explicit fma()
    "-fmad=false",
    "-Xcompiler", "-fPIC", "-shared",
    "-I", os.path.join(ROOT, "include"),
]


def nvcc_path() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libldrones.so cannot be built")


def is_stale() -> bool:
    if not (os.path.exists(LIB) and os.path.exists(LIB_CHECKED)):
        return True
    t = min(os.path.getmtime(LIB), os.path.getmtime(LIB_CHECKED))
    return any(os.path.getmtime(f) > t for f in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB
    checked = subprocess.Popen([nvcc_path()] + NVCC_FLAGS + ["-DLD_CHECKED", "-o", LIB_CHECKED] + SOURCES, stdout=subprocess.PIPE,
                               stderr=subprocess.STDOUT, text=True)  # in parallel with the product build
    cmd = [nvcc_path()] + NVCC_FLAGS + ["-Xptxas", "-v", "-o", LIB] + SOURCES
    r = subprocess.run(cmd, capture_output=True, text=True)
    checked_out, _ = checked.communicate()
    if checked.returncode != 0:
        sys.stderr.write(checked_out)
        raise RuntimeError("nvcc failed building libldrones_checked.so")
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libldrones.so")
    spilled = episode_kernel_spills(r.stdout + r.stderr)
    if spilled:
        os.remove(LIB)
        raise RuntimeError(f"the fp64 episode kernel spills registers ({spilled}): its hot loop must stay in registers")
    for lib in (LIB, LIB_CHECKED):
        bad = verify_tmem_scoreboards(lib)
        if bad:
            os.remove(lib)
            raise RuntimeError(f"{os.path.basename(lib)}: a tcgen05.ld result could be read before its scoreboard is waited for:\n  " + "\n  ".join(bad[:10]))
    return LIB


def episode_kernel_spills(ptxas_log: str):
    """ptxas -v lines of the fly_mlp_f64_kernel instantiations that report spill stores / loads (a spilling hot loop
    costs 30 % — measured)."""
    bad, current = [], None
    for line in ptxas_log.splitlines():
        if "Function properties for" in line or "Compiling entry function" in line:
            current = line
        elif "spill stores" in line and current and "fly_mlp_f64_kernel" in current and " 0 bytes spill stores, 0 bytes spill loads" not in line:
            bad.append(line.strip())
    return bad


# ---- SASS check of the episode kernel's Tensor-Memory loads ---------------------------------------------------------------
# The hot loop streams the brain with tcgen05.ld and WITHOUT a tcgen05.wait::ld between the loads (csrc/ldrones.cu,
# mlp_tmem_stream: the PTX-level wait is global — it waits for every outstanding load — and cost ~7 % of the kernel).  What
# makes that correct is the hardware scoreboard ptxas attaches to every LDTM and the wait it puts on the load's first
# consumer.  This check proves the property on the shipped SASS, for every instantiation of the episode kernel: walking
# forward from each LDTM along every path of the control-flow graph, no instruction may touch the load's destination
# registers before an instruction whose control word waits on the load's scoreboard (the wait is applied before that
# instruction reads its operands).  Control-word fields (sm_70+ encoding): bits 41-44 stall count, 46-48 write scoreboard, 52-57 wait mask.
_KNOWN_32BIT = {"IMAD", "IADD3", "LOP3", "SHF", "LEA", "ISETP", "FSEL", "SEL", "MOV", "VIADD", "VIMNMX3", "VIMNMX", "IABS", "PRMT", "R2UR",
                "S2R", "VOTE", "SHFL", "PLOP3", "NOP", "MUFU", "FMUL", "FADD", "FFMA", "FSETP", "I2FP", "F2I", "IMNMX", "POPC", "FLO", "BREV",
                "ISCADD", "LDC", "LDS", "STS", "LDG", "STG", "ATOMG", "ATOMS", "RED", "LDGSTS", "FMNMX", "I2F", "F2F", "FCHK", "SGXT", "BMSK",
                "P2R", "R2P", "MATCH", "REDUX", "LDL", "STL", "NANOSLEEP", "ERRBAR", "MEMBAR", "CCTL", "DEPBAR", "YIELD", "BAR", "WARPSYNC"}


def _sass_opcode(text):
    t = text.split()
    return (t[1] if t[0].startswith("@") else t[0]).split(".")[0]


def _sass_registers(text):
    """Registers an instruction may read or write, as a set of register numbers (over-approximated by operand width)."""
    import re

    op = _sass_opcode(text)
    full = text.split()[1] if text.split()[0].startswith("@") else text.split()[0]
    if op in ("LDTM", "STTM"):
        m = re.search(r"\.x(\d+)", full)
        width = int(m.group(1)) if m else 1
    elif ".128" in text:
        width = 4
    elif op.startswith("D") or any(k in text for k in (".64", "F64", "WIDE", "CS2R")):
        width = 2
    elif op in _KNOWN_32BIT:
        width = 1
    else:
        width = 2
    regs = set()
    for m in re.finditer(r"(?<![A-Za-z0-9_])R(\d+)\b", text):
        n = int(m.group(1))
        regs.update(range(n, n + width))
    return regs


def check_sass_tmem_loads(instructions):
    """instructions: list of (address, text, control_word_high) of ONE function.  Returns a list of violations (empty =
    proven).  Every path of the control-flow graph that leaves an LDTM is followed (conditional branches fork, loops are
    cut at revisits) until an instruction waits on the load's scoreboard."""
    import re

    index_of = {a: i for i, (a, _, _) in enumerate(instructions)}
    bad = []
    for k, (addr, text, w1) in enumerate(instructions):
        if _sass_opcode(text) != "LDTM":
            continue
        # scoreboard 7 = none: ptxas then relies on the in-order completion of Tensor-Memory loads and tracks the NEXT
        # LDTM, which must follow before anything touches this one's registers
        full = text.split()[1] if text.split()[0].startswith("@") else text.split()[0]
        m = re.search(r"\.x(\d+)", full)
        width = int(m.group(1)) if m else 1
        base = int(re.search(r"(?<![A-Za-z0-9_])R(\d+)\b", text).group(1))
        dest = set(range(base, base + width))
        stack, seen = [(k + 1, (w1 >> 46) & 7)], set()
        while stack and len(bad) < 50:
            i, sb = stack.pop()
            while True:
                if (i, sb) in seen:
                    break
                seen.add((i, sb))
                if i >= len(instructions):
                    bad.append(f"{addr:#x} {text}: its scoreboard ({sb}) is never waited for")
                    break
                a2, t2, c2 = instructions[i]
                op = _sass_opcode(t2)
                if sb == 7:
                    if op == "LDTM" and ((c2 >> 46) & 7) != 7 and not (_sass_registers(t2) & dest):
                        sb = (c2 >> 46) & 7
                        i += 1
                        continue
                elif (c2 >> 52) & 0x3F & (1 << sb):
                    break  # this instruction waits for the load before it reads anything: path proven
                if _sass_registers(t2) & dest:
                    bad.append(f"{addr:#x} {text}: {a2:#x} {t2} touches its registers before scoreboard ({sb}) is waited for")
                    break
                conditional = t2.startswith("@") or "UP" in t2.split(None, 1)[-1].split(",")[0]
                if op == "BRA":
                    target = [x.rstrip(",") for x in t2.split() if x.startswith("0x")]
                    if not target or int(target[-1], 16) not in index_of:
                        bad.append(f"{addr:#x} {text}: branch {a2:#x} {t2} with an unknown target while the load is in flight")
                        break
                    stack.append((index_of[int(target[-1], 16)], sb))
                    if not conditional:
                        break
                elif op in ("EXIT", "RET"):
                    if not conditional:
                        break  # the path ends without reading the registers
                elif op in ("BRX", "JMP", "JMX", "CALL", "BREAK", "KILL", "RPCMOV", "TRAP"):
                    bad.append(f"{addr:#x} {text}: {a2:#x} {t2} while the load is in flight")
                    break
                i += 1
    return bad


def parse_sass(listing):
    """cuobjdump -sass text of ONE function -> [(address, text, control_word_high)]"""
    import re

    ins, cur = [], None
    for ln in listing.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);\s+/\* 0x([0-9a-f]{16}) \*/", ln)
        if m:
            cur = [int(m.group(1), 16), m.group(2).strip(), None]
            ins.append(cur)
            continue
        m = re.match(r"\s+/\* 0x([0-9a-f]{16}) \*/", ln)
        if m and cur is not None and cur[2] is None:
            cur[2] = int(m.group(1), 16)
    return [tuple(i) for i in ins if i[2] is not None]


def verify_tmem_scoreboards(lib: str, pattern: str = "fly_mlp_f64_kernel"):
    cuobjdump = os.path.join(os.path.dirname(nvcc_path()), "cuobjdump")
    listing = subprocess.run([cuobjdump, "-sass", lib], capture_output=True, text=True, check=True).stdout
    bad, checked = [], 0
    for block in listing.split("Function : ")[1:]:
        name = block.split("\n", 1)[0].strip()
        if pattern not in name:
            continue
        ins = parse_sass(block)
        loads = sum(1 for _, t, _ in ins if _sass_opcode(t) == "LDTM")
        if loads == 0:
            bad.append(f"{name}: no LDTM found (parser out of date?)")
        checked += 1
        tag = name[name.find(pattern):][:len(pattern) + 24]
        bad += [f"{tag}: {b}" for b in check_sass_tmem_loads(ins)]
    if checked == 0:
        bad.append(f"no function matching {pattern} in {lib}")
    return bad


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
