"""Per-layer LoRA gradient comparison against the oracle (debug helper, run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import pcm_ref, unet_ref
from pcm_b200 import config
from pcm_b200.step import PCMTrainStep

B, hw, mp = 2, 16, 4
P = unet_ref.init_params(unet_ref.TINY, 0)
batch = pcm_ref.make_batch(unet_ref.TINY, B, hw, seed=0)
ref = pcm_ref.pcm_step_ref(unet_ref.TINY, P, batch, multiphase=mp, emulate_bf16=True)
st = PCMTrainStep(config.TINY, P, torch.device("cuda"), batch=B, height=hw, width=hw, multiphase=mp, keep_debug=True)
nhwc = lambda x: x.permute(0, 2, 3, 1).contiguous()
st.load_inputs(nhwc(batch["latents"]), nhwc(batch["noise"]), batch["index"], batch["w"],
               batch["prompt_embeds"].bfloat16(), batch["uncond_prompt_embeds"].bfloat16())
st.forward_backward()
torch.cuda.synchronize()
print("loss", st.loss.item(), ref["loss"].item())
g = st.unet.lora_grad_dict()
bad = 0
for k, rg in ref["grads"].items():
    gg = g[k].cpu().float()
    rel = ((gg - rg).norm() / (rg.norm() + 1e-20)).item()
    ratio = (gg.norm() / (rg.norm() + 1e-20)).item()
    if rel > 0.05:
        bad += 1
        print(f"{rel:9.3f} ratio {ratio:8.3f} |ref| {rg.norm().item():.3e}  {k}")
print("bad layers:", bad, "of", len(ref["grads"]))
