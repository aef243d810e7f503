"""Static complement to tools/ct_counts.py: inventory of control flow in the SASS of the secret-handling kernels (CPU only, cuobjdump).
For each kernel: uniform branches (BRA.U: the predicate lives in the uniform datapath, identical for every lane by construction),
per-lane predicated branches with the compare that feeds them, divergence regions (BSSY/BSYNC), indirect branches (BRX/JMX: none
expected), and loads whose address register is produced by a per-lane SEL/LOP3 chain are NOT analysed here -- the dynamic counters
(identical sector counts across secrets, profiles/r02e_ct_counts.txt) cover addresses.  The per-lane branches are listed so that a
reader can check each against the source: they are loop back-edges on counters and the public-data conditions (status != Ok early
exit; the rare carry ripple of fe_add_v / fe_sub_v inside fb_accumulate, whose operands are public).
usage: python tools/ct_sass_inventory.py > profiles/r02v_ct_sass_inventory.txt"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "anonymous-credit-tokens_b200", "libact_b200.so")
KERNELS = ["refund_sign_kernel", "refund_sign_seq_kernel", "issue_kernel", "issue_mode_kernel", "spend_head_kernel", "finalize_ctx_kernel", "public_key_kernel"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", sass)[1:]
    print(__doc__.split("usage:")[0])
    for f in funcs:
        name = f.split("\n", 1)[0].strip()
        short = next((k for k in KERNELS if re.search(r"\d+" + k + r"P", name)), None)
        if not short:
            continue
        ins = re.findall(r"/\*([0-9a-f]{4,6})\*/\s+([^;]+);", f)
        ops = [(int(a, 16), t.strip()) for a, t in ins]
        n = len(ops)
        bra_u = sum(1 for _, t in ops if re.match(r"(@!?UP\d\s+)?BRA\.U", t) or ("BRA" in t and "UP" in t.split("BRA")[0]))
        lane = [(a, t) for a, t in ops if re.match(r"@!?P\d\s+BRA", t)]
        uncond = sum(1 for _, t in ops if t.startswith("BRA ") )
        bssy = sum(1 for _, t in ops if t.startswith("BSSY"))
        indirect = [t for _, t in ops if t.startswith(("BRX", "JMX", "JMP", "BRXU"))]
        calls = sum(1 for _, t in ops if t.startswith("CALL"))
        print(f"## {short}: {n} SASS instructions, {calls} CALL sites, {bra_u} uniform branches, {uncond} unconditional, {len(lane)} per-lane predicated branches, "
              f"{bssy} divergence regions (BSSY), {len(indirect)} indirect branches")
        for a, t in lane:
            # nearest preceding instruction that writes this predicate
            m = re.match(r"@!?(P\d)", t)
            p = m.group(1)
            src = next((u for b, u in reversed([o for o in ops if o[0] < a]) if re.search(r"\b" + p + r"\b", u.split(",")[0]) and not u.startswith("@")), "?")
            back = "back-edge" if (re.search(r"0x([0-9a-f]+)", t) and int(re.search(r"0x([0-9a-f]+)", t).group(1), 16) < a) else "forward"
            print(f"    {a:#07x}  {t:40s} <- {src[:90]}   [{back}]")
        print()


if __name__ == "__main__":
    main()
