"""Where the module path's 76 ms per token go: host-side cProfile of one HF decode step (BLOCKS layers) with swapped modules."""
import cProfile, pstats, os, sys, time, io
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("BLOCKS", "8")
os.environ["STEPS"] = "8"
src = open(os.path.join(os.path.dirname(__file__), "module_path_speed.py")).read()
src = src[:src.index("torch.cuda.synchronize()\nids = ")]
exec(src)
ids = torch.randint(0, shape.vocab - 1, (1, 64), device=dev)
with torch.inference_mode():
    out = model(ids, use_cache=True)
    past, tok = out.past_key_values, out.logits[:, -1:].argmax(-1)
    for _ in range(8):
        out = model(tok, past_key_values=past, use_cache=True); past, tok = out.past_key_values, out.logits[:, -1:].argmax(-1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(16):
        out = model(tok, past_key_values=past, use_cache=True); past, tok = out.past_key_values, out.logits[:, -1:].argmax(-1)
    torch.cuda.synchronize()
    print("ms per token (", nb, "blocks):", (time.perf_counter() - t0) / 16 * 1e3)
    pr = cProfile.Profile(); pr.enable()
    for _ in range(16):
        out = model(tok, past_key_values=past, use_cache=True); past, tok = out.past_key_values, out.logits[:, -1:].argmax(-1)
    torch.cuda.synchronize()
    pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(25); print(s.getvalue()[:6000])
