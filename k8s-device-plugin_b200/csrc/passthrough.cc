// passthrough.cc — the rest of the reference hook's exported surface: 412 driver-API / NVML entry points that
// lib/nvidia/libvgpu.so wraps only to log and forward ("each = log + ENSURE_RUNNING + call
// through cuda_library_entry[i].fn"). Programs that link the hook's symbols directly (instead of going through
// dlsym/cuGetProcAddress, which already fall through to the real library for unhooked names) keep resolving.
// Each entry is a signature-agnostic trampoline: load the cached real address, tail-jump; the first call resolves
// it through the real dlsym with every argument register preserved. No per-call logging, locking or getenv.
// NOTE: cuda.h / nvml.h must NOT be included here: their version macros (cuGraphInstantiate -> ..WithFlags,
// nvmlDeviceGetCount -> .._v2) would rewrite the exported names below.
namespace vgpu {
void *real_cuda_symbol(const char *name);
void *real_nvml_symbol(const char *name);
}  // namespace vgpu

namespace {
struct Slot { void *fn; const char *name; int lib; };   // lib: 0 = libcuda, 1 = libnvidia-ml
extern "C" int vgpu_pass_missing_cuda() { return 500; }   // CUDA_ERROR_NOT_FOUND
extern "C" int vgpu_pass_missing_nvml() { return 13; }    // NVML_ERROR_FUNCTION_NOT_FOUND
}  // namespace

extern "C" __attribute__((visibility("hidden"), used)) void *vgpu_pass_resolve(Slot *s) {
    void *p = s->lib ? vgpu::real_nvml_symbol(s->name) : vgpu::real_cuda_symbol(s->name);
    if (!p) p = s->lib ? reinterpret_cast<void *>(&vgpu_pass_missing_nvml) : reinterpret_cast<void *>(&vgpu_pass_missing_cuda);
    __atomic_store_n(&s->fn, p, __ATOMIC_RELEASE);
    return p;
}

// r11 = slot. Preserve the integer and vector argument registers (and rax: vararg vector count), resolve, jump.
asm(R"(
    .text
    .type vgpu_pass_thunk,@function
vgpu_pass_thunk:
    pushq %rbp
    movq  %rsp, %rbp
    subq  $192, %rsp
    andq  $-16, %rsp
    movq  %rdi, 0(%rsp)
    movq  %rsi, 8(%rsp)
    movq  %rdx, 16(%rsp)
    movq  %rcx, 24(%rsp)
    movq  %r8,  32(%rsp)
    movq  %r9,  40(%rsp)
    movq  %rax, 48(%rsp)
    movdqu %xmm0, 64(%rsp)
    movdqu %xmm1, 80(%rsp)
    movdqu %xmm2, 96(%rsp)
    movdqu %xmm3, 112(%rsp)
    movdqu %xmm4, 128(%rsp)
    movdqu %xmm5, 144(%rsp)
    movdqu %xmm6, 160(%rsp)
    movdqu %xmm7, 176(%rsp)
    movq  %r11, %rdi
    call  vgpu_pass_resolve
    movq  %rax, %r11
    movq  0(%rsp), %rdi
    movq  8(%rsp), %rsi
    movq  16(%rsp), %rdx
    movq  24(%rsp), %rcx
    movq  32(%rsp), %r8
    movq  40(%rsp), %r9
    movq  48(%rsp), %rax
    movdqu 64(%rsp), %xmm0
    movdqu 80(%rsp), %xmm1
    movdqu 96(%rsp), %xmm2
    movdqu 112(%rsp), %xmm3
    movdqu 128(%rsp), %xmm4
    movdqu 144(%rsp), %xmm5
    movdqu 160(%rsp), %xmm6
    movdqu 176(%rsp), %xmm7
    leave
    jmp   *%r11
    .size vgpu_pass_thunk, .-vgpu_pass_thunk
)");

#define PASS_IMPL(name, libid)                                                                     \
    extern "C" { __attribute__((visibility("hidden"), used)) Slot vgpu_slot_##name = {nullptr, #name, libid}; } \
    asm(".text\n .globl " #name "\n .type " #name ",@function\n" #name ":\n"                       \
        "  movq vgpu_slot_" #name "(%rip), %r11\n  testq %r11, %r11\n  jnz 1f\n"                   \
        "  leaq vgpu_slot_" #name "(%rip), %r11\n  jmp vgpu_pass_thunk\n1: jmp *%r11\n"            \
        " .size " #name ", .-" #name "\n");
#define PASS_CU(name) PASS_IMPL(name, 0)
#define PASS_NVML(name) PASS_IMPL(name, 1)
#include "passthrough_names.inc"
