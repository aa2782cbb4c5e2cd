import os, sys
os.environ["KB200_CONV_F16"] = "1"; os.environ["KB_GRAPHS"] = "0"; os.environ.setdefault("KB200_RANDOM_VGG", "1")
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import torch, kb_helpers
torch.set_grad_enabled(False)
which = sys.argv[1]
H, W = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (192, 256)
dev = "cuda"
if which == "partial":
    from ken_burns_effect_b200.models.partial_inpainting import Inpaint as Net
    net = kb_helpers.deterministic_state(Net().eval()).to(dev)
    data = torch.randn(1, 68, H, W, device=dev); mask = (torch.rand(1, 1, H, W, device=dev) > 0.2).float()
    net.normalize_images_disp(torch.rand(1, 3, H, W, device=dev), torch.rand(1, 1, H, W, device=dev), True)
    o = net(mask, tensorData=data); torch.cuda.synchronize(); print(which, "ok", float(o['tensorImage'].mean()))
elif which == "refine":
    from ken_burns_effect_b200.models.disparity_refinement import Refine
    net = kb_helpers.deterministic_state(Refine().eval()).to(dev)
    o = net(torch.rand(1, 3, H, W, device=dev), torch.rand(1, 1, H // 4, W // 4, device=dev)); torch.cuda.synchronize(); print(which, "ok", float(o.mean()))
elif which == "disparity":
    from ken_burns_effect_b200.models.disparity_estimation import Disparity, Semantics
    sem = kb_helpers.deterministic_state(Semantics().eval()).to(dev); dis = kb_helpers.deterministic_state(Disparity().eval()).to(dev)
    x = torch.rand(1, 3, H, W, device=dev); o = dis(x, sem(x)); torch.cuda.synchronize(); print(which, "ok", float(o.mean()))
elif which == "inpaint":
    from ken_burns_effect_b200.models.pointcloud_inpainting import Inpaint as Net
    net = kb_helpers.deterministic_state(Net().eval()).to(dev)
    data = torch.randn(1, 68, H, W, device=dev); mask = (torch.rand(1, 1, H, W, device=dev) > 0.2).float()
    net.normalize_images_disp(torch.rand(1, 3, H, W, device=dev), torch.rand(1, 1, H, W, device=dev), True)
    o = net(mask, tensorData=data); torch.cuda.synchronize(); print(which, "ok", float(o['tensorImage'].mean()))
if which == "refine_steps":
    from ken_burns_effect_b200.models.disparity_refinement import Refine
    from ken_burns_effect_b200.utils import convstack as cs
    net = kb_helpers.deterministic_state(Refine().eval()).to(dev)
    orig = cs.conv2d
    n = [0]

    def traced(x, pc, outs, **kw):
        n[0] += 1
        print(f"conv {n[0]}: x {tuple(x.shape)} {x.dtype} stride {x.stride(2)} -> Cout {pc.Cout} k{pc.k} s{pc.stride} outs "
              f"{[(o[1], None if o[2] is None else (tuple(o[2].shape), o[2].dtype, o[2].stride(2))) for o in outs]} kw {list(kw)}", flush=True)
        r = orig(x, pc, outs, **kw)
        torch.cuda.synchronize()
        print("   done", flush=True)
        return r
    cs.conv2d = traced
    o = net(torch.rand(1, 3, H, W, device=dev), torch.rand(1, 1, H // 4, W // 4, device=dev)); torch.cuda.synchronize(); print(which, "ok")
