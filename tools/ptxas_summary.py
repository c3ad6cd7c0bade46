"""profiles/r01_ptxas_registers.txt from the per-file `-Xptxas -v` logs the Makefile writes (csrc/*.ptxas.log)."""
import glob, os, re, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = ["# ptxas -v summary (sm_100a, nvcc 12.9): registers / spills per kernel entry",
       "# regenerate: make -C thunder_speech_b200/csrc && python tools/ptxas_summary.py", ""]
pat = re.compile(r"Compiling entry function '([^']+)' for 'sm_100a'\n(?:ptxas info\s*: Function properties for [^\n]+\n)?"
                 r"\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\nptxas info\s*: Used (\d+) registers")
for f in sorted(glob.glob(os.path.join(ROOT, "thunder_speech_b200", "csrc", "*.ptxas.log"))):
    ents = pat.findall(open(f).read())
    if not ents:
        continue
    out.append("## " + os.path.basename(f).replace(".ptxas.log", ".cu"))
    for e in ents:
        name = subprocess.run(["c++filt", e[0]], capture_output=True, text=True).stdout.strip()
        out.append(f"{int(e[4]):4d} regs  spill {int(e[2]):3d}/{int(e[3]):3d} B  {re.sub(r'[(].*', '', name)[:90]}")
    out.append("")
open(os.path.join(ROOT, "profiles", "r01_ptxas_registers.txt"), "w").write("\n".join(out))
