"""GPU probe: ViT-B/16 DyT video model (BASELINE configs[4] shape per GPU: 8 clips x 8 frames x 224^2)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dynamic-tuning_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import dyt_oracle as O
from video_models.video_vision_transformer_IN21K import vit_base_patch16_224_in21k
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import gen_golden as G
dev = torch.device("cuda:0")
class Cfg(dict):
    __getattr__ = dict.__getitem__
tuning = Cfg(ffn_adapt=True, ffn_option="parallel", ffn_adapter_layernorm_option="none", ffn_adapter_init_option="lora",
             ffn_adapter_scalar="0.1", ffn_num=64, d_model=768, vpt_on=False, vpt_num=0)
select = Cfg(open=True, keep_layers=0, token_target_ratio=0.5)
sd = G.video_state_dict(O.synthetic_state_dict(seed=0, num_classes=174), 768, 0)
m = vit_base_patch16_224_in21k(num_classes=174, tuning_config=tuning, select_config=select)
m.load_state_dict(sd, strict=True)
m = m.eval().to(dev)
clips = int(os.environ.get("CLIPS", "8"))
x = torch.randn(clips, 3, 8, 224, 224, generator=torch.Generator().manual_seed(0)).to(dev)
def fwd():
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        return m(x)
for _ in range(3): out = fwd()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); a.record()
n = 10
for _ in range(n): out = fwd()
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / n
print(f"video: {clips} clips x 8 frames: {ms:.2f} ms/step, {clips / ms * 1e3:.1f} clips/s, keep {float(out[1]['token_select'].float().mean()):.3f}, logits {tuple(out[0].shape)}")
