"""CPU oracle for the AVT hot path (TEST INFRASTRUCTURE — never imported by avt_b200/).

A plain-PyTorch (fp32/fp64, CPU) restatement of what the reference computes on the path named in
BASELINE.json: timm-0.4.12 ViT (un-vendored third party), HF transformers-4.2.2 GPT-2 (un-vendored third
party) and the reference's own AVTh / BaseModel glue (models/future_prediction.py, models/base_model.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package, and only as the checker or the reported CPU baseline.

Pinning: the reference ships no tests or golden vectors, so the oracle is pinned against
  (1) the UNMODIFIED reference classes (BaseModel, TIMMModel, AVTh) executed in the authoring container
      from /root/reference under stub modules (oracle/ref_host.py) — outputs and gradients committed as
      fixtures under tests/golden/ by oracle/gen_golden.py;
  (2) torchvision.models.VisionTransformer (same math as timm's ViT, different key names);
  (3) the installed transformers GPT2Model (5.5.0; same math as 4.2.2).
"""
