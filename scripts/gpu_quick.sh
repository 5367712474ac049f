#!/bin/bash
python - <<'PY'
import sys, json, torch
sys.path.insert(0, ".")
from cacophony_b200 import _lib as L, ops
B,S,H,dh=256,500,8,96
qkv=(torch.randn(B,S,3*H*dh,device="cuda")).half(); mask=torch.ones(B,S,device="cuda"); mask[:,496:]=0
lib=L.load(); lib.caco_set_attention_impl(4)
def t():
    for _ in range(3): ops.attention_audio(qkv,mask,H)
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): ops.attention_audio(qkv,mask,H)
    e1.record(); torch.cuda.synchronize()
    return round(e0.elapsed_time(e1)/20,4)
res={"spec":[], "two_pass":[], "old":[]}
for r in range(4):
    lib.caco_attn3_set_speculative(1); res["spec"].append(t())
    lib.caco_attn3_set_speculative(0); res["two_pass"].append(t())
    lib.caco_set_attention_impl(6); res["old"].append(t()); lib.caco_set_attention_impl(4)
print(json.dumps(res))
PY
