# Diagnostic script (not collected by pytest): lives under tests/ because it uses the oracle as the checker.
import sys; sys.path.insert(0, '.')
import numpy as np, torch
import oracle, two_tower_b200 as tt
from two_tower_b200 import synth
def rel(a,b):
    a=np.asarray(a,np.float64); b=np.asarray(b,np.float64); return float(np.abs(a-b).max()/max(np.abs(b).max(),1e-30))
def rms(a,b):
    a=np.asarray(a,np.float64); b=np.asarray(b,np.float64); return float(np.sqrt(((a-b)**2).mean())/np.sqrt((b**2).mean()))
tt.set_precision("bf16")
vu, vi, d, mlp, B, T = 2000, 1500, 64, (128, 64), 512, 0.5
um = tt.Sequential([tt.layers.Embedding(vu, d), tt.layers.Dense(mlp[0], "relu"), tt.layers.Dense(mlp[1])])
im = tt.Sequential([tt.layers.Embedding(vi, d), tt.layers.Dense(mlp[0], "relu"), tt.layers.Dense(mlp[1])])
task = tt.tasks.Retrieval(temperature=T)
rng = synth.rng_for(31)
b = {"u": synth.draw_ids(rng, B, vu, 1.3), "i": synth.draw_ids(rng, B, vi, 1.3)}
um(b["u"]); im(b["i"])
get = lambda seq, name: {"tables": {name: seq.layers[0].get_weights()[0].astype(np.float64)},
                         "kernels": [l.get_weights()[0].astype(np.float64) for l in seq.layers[1:]],
                         "biases": [l.get_weights()[1].astype(np.float64) for l in seq.layers[1:]]}
qs = oracle.TowerSpec([("u", "id", vu, None)], d, mlp); cs = oracle.TowerSpec([("i", "id", vi, None)], d, mlp)
qp, cp = get(um, "u"), get(im, "i")
with tt.GradientTape() as tape:
    q = um(b["u"]); c = im(b["i"]); loss = task(q, c)
    vs = um.trainable_variables + im.trainable_variables
    grads = tape.gradient(loss, vs)
for mode in (True, False):
    q_ref, qc = oracle.tower_forward(qs, qp, {"u": b["u"]}, bf16=mode)
    c_ref, cc = oracle.tower_forward(cs, cp, {"i": b["i"]}, bf16=mode)
    if mode: q_ref, c_ref = oracle.bf16_round(q_ref).astype(np.float64), oracle.bf16_round(c_ref).astype(np.float64)
    r = oracle.retrieval_loss_and_grads(q_ref, c_ref, temperature=T)
    print("oracle bf16-emulation" if mode else "oracle fp64", "loss", loss.item(), r["loss"])
    print("  q", rel(q.numpy(), q_ref), rms(q.numpy(), q_ref))
    print("  dq", rel(q.grad["f32"].cpu().numpy(), r["dq"]), rms(q.grad["f32"].cpu().numpy(), r["dq"]))
    print("  dc", rel(c.grad["f32"].cpu().numpy(), r["dc"]), rms(c.grad["f32"].cpu().numpy(), r["dc"]))
    dk, db, sp = oracle.tower_backward(qs, qp, {"u": b["u"]}, qc, r["dq"])
    g = {v.name: gg for v, gg in zip(vs, grads)}
    print("  dx", rel(g[um.layers[0].embeddings.name].rows.cpu().numpy(), sp["u"][1]), rms(g[um.layers[0].embeddings.name].rows.cpu().numpy(), sp["u"][1]))
    for j, l in enumerate(um.layers[1:]):
        print("  dk%d"%j, rel(g[l.kernel.name].parts.sum(0).cpu().numpy(), dk[j]), rms(g[l.kernel.name].parts.sum(0).cpu().numpy(), dk[j]),
              " db%d"%j, rel(g[l.bias.name].parts.sum(0).cpu().numpy().reshape(-1), db[j]))
    # feed the oracle's dq through my dense bwd to isolate the MLP backward
