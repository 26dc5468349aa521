"""Does a Triton-compiled TMA (tensor descriptor) load run on this box?"""
import torch, triton, triton.language as tl
print("triton", triton.__version__)
try:
    from triton.tools.tensor_descriptor import TensorDescriptor

    @triton.jit
    def k(in_desc, out_ptr, M: tl.constexpr, N: tl.constexpr):
        x = in_desc.load([8, 32])
        offs = tl.arange(0, M)[:, None] * N + tl.arange(0, N)[None, :]
        tl.store(out_ptr + offs, x)

    a = torch.arange(256 * 256, device="cuda", dtype=torch.float32).reshape(256, 256)
    desc = TensorDescriptor.from_tensor(a, [16, 32])
    out = torch.empty(16, 32, device="cuda")
    h = k[(1,)](desc, out, 16, 32)
    torch.cuda.synchronize()
    print("triton TMA load ok:", bool((out == a[8:24, 32:64]).all()))
    try:
        sass = h.asm.get("sass", "") if hasattr(h, "asm") else ""
        print("UTMALDG in triton sass:", "UTMALDG" in sass)
        ptx = h.asm.get("ptx", "")
        print([l.strip() for l in ptx.splitlines() if "cp.async.bulk.tensor" in l][:3])
    except Exception as e:
        print("asm inspect failed", e)
except Exception as e:
    import traceback; traceback.print_exc()
