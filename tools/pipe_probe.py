"""FP32 pipe probes: packed-op throughput of FFMA/FFMA2/FADD2/FMUL2 and of FFMA2 mixed with MUFU.RSQ.
Numbers are reported as 'FMA-equivalent TFLOP/s' (2 flop per lane-op) so they compare directly."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "lambda-cdm-raytracing_b200", "python"))
import b200grav
e = b200grav.Engine(0)
names = {0: "FFMA", 1: "FFMA2", 2: "FADD2", 3: "FMUL2", 4: "FFMA2 + 1.25 MUFU.RSQ per 8", 5: "FFMA2 + 2 MUFU.RSQ per 8",
         6: "DFMA only (8 chains; counted as if FFMA2)", 7: "FFMA2 + DFMA interleaved 1:1 (FFMA2 count)",
         8: "FFMA, 3 distinct register operands", 9: "FFMA2, 3 distinct register-pair operands",
         10: "FFMA2, 2 distinct register-pair operands"}
for m in range(11):
    t, ms = e.fp32_peak_probe(m, 2000)
    print(f"mode {m} {names[m]:32s} {t:7.2f} TFLOP/s-equivalent  ({ms:.2f} ms)")
