import json,sys
d=json.load(open(sys.argv[1])); k=d.pop("kernels")
print(sys.argv[1], round(d["value"]), "pairs/s", round(d["ms_per_step"],3), "ms; e2e", round(d["e2e"]["value"]), "gemm", round(d["gemm_summary"]["ms_per_step"],2), "ms", round(d["gemm_summary"]["tflops"]), "TF; kernels", round(d["kernel_ms_per_step"],2), "launches", d["gpu_launches"])
print("   roofline:", {x: d["roofline"].get(x) for x in ("kernel", "bound", "frac")})
for n,v in list(k.items())[:int(sys.argv[2]) if len(sys.argv)>2 else 8]: print("   ", n, round(v["ms_per_step"],3), round(v.get("tflops",0)), "TF", round(v.get("gbs",0)), "GB/s", v.get("bound"), round(v.get("roofline_frac",0),2))
