"""Build the REFERENCE's own CUDA kernels for sm_100a, in place from /root/reference, into
oracle/_ref/ (git-ignored; travels to the GPU box with the gpurun snapshot).  They are the GPU-side
"kernel to beat" (SURVEY §11-6): auto_gptq.vecquant{2,3,4}matmul_faster_old and
faster_transformer.gemv_4bit / gemm_4bit.  Nothing is copied into the repo: the reference sources are
compiled where they lie; the only file written here is a 10-line pybind shim for the two FT ops
(the reference's FT.cpp also pulls in attention / layernorm, which are out of scope).

    python oracle/build_ref_kernels.py          # ~3 min per translation unit (torch headers)
"""
import os
import sys

REF = "/root/reference/amq/kernel"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")


def main():
    if not os.path.isdir(REF):
        print("reference not present; nothing to build")
        return 0
    os.makedirs(OUT, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    os.environ.setdefault("MAX_JOBS", "4")
    from torch.utils.cpp_extension import load
    flags = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-U__CUDA_NO_HALF_OPERATORS__",
             "-U__CUDA_NO_HALF_CONVERSIONS__", "-U__CUDA_NO_HALF2_OPERATORS__", "--expt-relaxed-constexpr"]
    which = sys.argv[1:] or ["auto_gptq", "ft"]
    if "auto_gptq" in which:
        load(name="auto_gptq", sources=[os.path.join(REF, "AutoGPTQ", "auto_gptq_kernel.cu")], extra_cuda_cflags=flags,
             build_directory=OUT, verbose=True, is_python_module=False, with_cuda=True)
    if "ft" in which:
        shim = os.path.join(OUT, "ft_quant_shim.cpp")
        with open(shim, "w") as f:
            f.write('#include <torch/extension.h>\n'
                    f'#include "{REF}/ft/quantization_new/gemv/gemv_cuda.h"\n'
                    f'#include "{REF}/ft/quantization_new/gemm/gemm_cuda.h"\n'
                    'PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {\n'
                    '  m.def("gemv_4bit", &gemv_4bit, "reference W4A16 GEMV");\n'
                    '  m.def("gemm_4bit", &gemm_4bit, "reference W4A16 GEMM");\n}\n')
        load(name="ft_quant_ref", sources=[shim, os.path.join(REF, "ft", "quantization_new", "gemv", "gemv_cuda.cu"),
                                           os.path.join(REF, "ft", "quantization_new", "gemm", "gemm_cuda.cu")],
             extra_cuda_cflags=flags + ["--use_fast_math"], extra_include_paths=[os.path.join(REF, "ft")],
             build_directory=OUT, verbose=True, is_python_module=False, with_cuda=True)
    print("built:", sorted(f for f in os.listdir(OUT) if f.endswith(".so")))
    return 0


if __name__ == "__main__":
    sys.exit(main())
