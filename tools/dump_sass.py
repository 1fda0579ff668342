"""Write the SASS evidence DESIGN.md quotes into profiles/ (runs anywhere: cuobjdump needs no GPU).

    python tools/dump_sass.py [round_tag]          (default r02)

For each kernel named below: the full `cuobjdump -sass` text of the function as built into
traceit_b200/libtraceit_b200.so, preceded by an opcode histogram, so a reader can check claims such as
"FFMA2 / UBLKCP / SYNCS in gravity_kernel", "STG.E.ENL2.256 in scatter_kernel", "no MEMBAR in the hand-off
of response_flow_kernel<1>" without rebuilding anything."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
SO = os.path.join(ROOT, "traceit_b200", "libtraceit_b200.so")
KERNELS = {   # file stem -> regex on the demangled function name
    "gravity_kernel": r"tr::gravity_kernel\(",
    "keys_kernel": r"tr::keys_kernel\(",
    "scatter_kernel": r"tr::scatter_kernel\(",
    "narrow_kernel_spheres": r"tr::narrow_kernel<\(bool\)0, \(bool\)0, \(bool\)0>\(",
    "response_flow_kernel_1lane": r"tr::response_flow_kernel<\(int\)1>\(",
    "prep_kernel": r"tr::prep_kernel\(",
    "sparse_step_kernel": r"tr::sparse_step_kernel\(",
    "tiny_step_kernel": r"tr::tiny_step_kernel\(",
}

text = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
demangle = subprocess.run(["cu++filt"], input=text, capture_output=True, text=True, check=True).stdout
blocks = re.split(r"\n(?=\s*Function : )", demangle)
for stem, pat in KERNELS.items():
    hit = [b for b in blocks if re.search(r"Function : .*" + pat, b.split("\n", 1)[0])]
    if not hit:
        print(f"{stem}: not found", file=sys.stderr)
        continue
    body = hit[0]
    ops = collections.Counter()
    for ln in body.splitlines():
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", ln)
        if m:
            ops[m.group(1)] += 1
    out = os.path.join(ROOT, "profiles", f"{tag}_sass_{stem}.txt")
    with open(out, "w") as f:
        f.write(f"# cuobjdump -sass traceit_b200/libtraceit_b200.so | cu++filt   (function matching /{pat}/)\n")
        f.write(f"# {sum(ops.values())} instructions; opcode histogram (top 40):\n")
        for op, c in ops.most_common(40):
            f.write(f"#   {c:6d}  {op}\n")
        # keep the instruction text, drop the hex encodings (second line of every instruction and the trailing column)
        for ln in body.splitlines():
            if re.match(r"\s*/\* 0x[0-9a-f]{16} \*/\s*$", ln):
                continue
            f.write(re.sub(r"\s*/\* 0x[0-9a-f]{16} \*/\s*$", "", ln).rstrip() + "\n")
    print(f"{out}: {sum(ops.values())} instructions")
