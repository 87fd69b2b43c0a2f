"""TEST INFRASTRUCTURE: builds a host-emulated copy of selected SIMT translation units of
msmdfusion_b200/csrc (see cuda_emul.h).  The CUDA sources are used as they are: the script takes the
device helpers of common.cuh and the whole body of the .cu file, rewrites the ``kernel<<<grid, block,
smem, stream>>>(args)`` launches into ``emu::launch(grid, block, [&]{ kernel(args); })`` and compiles the
result with g++ -std=c++20.  Entry points keep their C-ABI names with an ``emu_`` prefix."""
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(HERE)))
CSRC = os.path.join(ROOT, 'msmdfusion_b200', 'csrc')
OUT = os.path.join(HERE, '_build')


def _device_helpers():
    """common.cuh from the '// device helpers' banner to the end of namespace msmd."""
    text = open(os.path.join(CSRC, 'common.cuh')).read()
    start = text.index('// device helpers')
    start = text.index('\n', text.index('// ----', start)) + 1
    end = text.rindex('}  // namespace msmd')
    ws0 = text.index('// Bump allocator')
    ws1 = text.index('// ----', ws0)
    return 'namespace msmd {\n' + text[ws0:ws1] + text[start:end] + '}  // namespace msmd\n'


def _rewrite_launches(src):
    out, pos = [], 0
    for m in re.finditer(r'([A-Za-z_]\w*(?:<[^<>;(){}]*>)?)\s*<<<', src):
        if m.start() < pos:
            continue
        cfg_end = src.index('>>>', m.end())
        cfg = src[m.end():cfg_end]
        # split the launch configuration at top-level commas
        parts, depth, cur = [], 0, ''
        for ch in cfg:
            if ch in '(<[':
                depth += 1
            elif ch in ')>]':
                depth -= 1
            if ch == ',' and depth == 0:
                parts.append(cur)
                cur = ''
            else:
                cur += ch
        parts.append(cur)
        p = src.index('(', cfg_end)
        depth, q = 0, p
        while True:
            if src[q] == '(':
                depth += 1
            elif src[q] == ')':
                depth -= 1
                if depth == 0:
                    break
            q += 1
        args = src[p + 1:q]
        out.append(src[pos:m.start()])
        smem = parts[2].strip() if len(parts) > 2 else '0'
        out.append(f'::emu::launch(dim3({parts[0].strip()}), dim3({parts[1].strip()}), (size_t)({smem}), [&]() {{ '
                   f'{m.group(1)}({args}); }})')
        pos = q + 1
    out.append(src[pos:])
    return ''.join(out)


_INLINED = set()


def _inline_headers(src):
    """#include "x.cuh" -> the header's text, once per unit (common.cuh is provided by cuda_emul.h +
    _device_helpers)."""
    def sub(m):
        name = m.group(1)
        if name == 'common.cuh' or name in _INLINED:
            return ''
        _INLINED.add(name)
        text = open(os.path.join(CSRC, name)).read().replace('#pragma once', '')
        return _inline_headers(text)
    return re.sub(r'#include "(\w+\.cuh)"', sub, src)


def translate(cu_name):
    src = _inline_headers(open(os.path.join(CSRC, cu_name)).read())
    src = _rewrite_launches(src)
    defined = re.findall(r'extern "C" MSMD_API \w[\w ]*?[ *]msmd_(\w+)\(', src)
    src = re.sub(r'extern "C" MSMD_API (\w[\w ]*?[ *])msmd_(\w+)\(', r'extern "C" \1emu_msmd_\2(', src)
    for name in defined:  # calls between entry points of this unit follow the renaming
        src = re.sub(r'(?<![\w])msmd_%s\(' % name, 'emu_msmd_%s(' % name, src)
    # calls between entry points of the same unit keep working; calls into OTHER units are stubbed
    return src


PRELUDE = '''#define MSMD_EMUL 1
#include "cuda_emul.h"
namespace emu {
thread_local dim3 t_threadIdx, t_blockIdx;
dim3 g_blockDim, g_gridDim;
thread_local WarpCtx* t_warp = nullptr;
thread_local int t_lane = 0;
BlockSync g_bsync;
std::function<void()>* g_body = nullptr;
char g_error[512];
std::atomic<int> g_or{0};
uint8_t* g_dyn_smem = nullptr;
void (*g_block_begin)(uint32_t) = nullptr;
void (*g_block_end)() = nullptr;
bool g_blocks_descending = false;
}
extern "C" const char* emu_last_error() { return ::emu::g_error; }
'''

# entry points of other translation units that spconv_bwd.cu calls (the forward kernels: tensor-core
# code cannot be emulated; dgrad-through-forward is covered by the oracle-level identity test instead)
STUBS_BWD = '''
extern "C" int msmd_spconv_fwd_tc_ws(const float*, int, const float*, const int*, int, int, int, int, const float*,
                                     const float*, const float*, int, float*, void*, size_t, msmd_stream_t) { return -100; }
extern "C" int msmd_spconv_fwd_tc16(const float*, int, const void*, const int*, const int*, int, int, int, int, int,
                                    const float*, const float*, const float*, int, float*, msmd_stream_t) { return -100; }
extern "C" int msmd_spconv_bwd_weight_tc_supported(int, int, int) { return 0; }
extern "C" size_t msmd_spconv_bwd_weight_tc_workspace(int, int, int, int) { return 0; }
extern "C" int msmd_spconv_bwd_weight_tc(const float*, int, const float*, const int*, int, int, int, int, float*, void*,
                                         size_t, msmd_stream_t) { return -100; }
extern "C" int msmd_spconv_fwd(const float*, int, const float*, const int*, int, int, int, int, const float*,
                               const float*, const float*, int, float*, msmd_stream_t);
extern "C" size_t msmd_grid_num_words(int batch_size, const int* s) {   // as csrc/rulebook.cu
  return (size_t)(((long long)batch_size * s[0] * s[1] * s[2] + 31) / 32);
}
'''


def build(verbose=False):
    os.makedirs(OUT, exist_ok=True)
    lib = os.path.join(OUT, 'libmsmd_emul.so')
    deps = [os.path.join(CSRC, f) for f in ('common.cuh', 'spconv_bwd.cu', 'spconv.cu', 'fusion.cu', 'sort.cuh',
                                            'scan.cuh')] + \
        [os.path.join(HERE, 'cuda_emul.h'), os.path.abspath(__file__)]
    if os.path.exists(lib) and all(os.path.getmtime(lib) > os.path.getmtime(d) for d in deps):
        return lib
    fwd = translate('spconv.cu').replace('emu_msmd_spconv_fwd(', 'msmd_spconv_fwd(')  # linked by bwd_data
    _INLINED.clear()
    text = PRELUDE + _device_helpers() + STUBS_BWD + fwd + translate('spconv_bwd.cu') + translate('fusion.cu')
    cpp = os.path.join(OUT, 'emul_unit.cpp')
    with open(cpp, 'w') as f:
        f.write(text)
    cmd = ['g++', '-std=c++20', '-O1', '-g', '-ffp-contract=off', '-shared', '-fPIC', '-pthread', '-I', HERE, cpp, '-o', lib]
    if verbose:
        print(' '.join(cmd))
    subprocess.check_call(cmd)
    return lib


TC_PRELUDE = '''#include "tc_emul.h"
namespace emu {
uint8_t* g_smem_window = nullptr;
uint32_t g_dyn_bytes = 0;
std::mutex g_tc_mu;
std::condition_variable g_tc_cv;
std::map<uint32_t, MBar> g_mbar;
uint32_t g_tmem[128][512];
uint32_t g_tmem_next = 0, g_tmem_live = 0;
NamedBar g_named[16];
std::atomic<long long> g_mma_count{0};
std::deque<AsyncOp> g_mma_queue;
std::vector<AsyncOp> g_copy_queue;
std::map<int, std::vector<std::function<void()>>> g_cpasync_pending;
int g_async_late = 1;
static void tc_begin(uint32_t bytes) { tc_block_reset(bytes); g_dyn_smem = g_smem_window + kDynBase; }
static void tc_end() { tc_block_check(); }
static struct TcHooks { TcHooks() { g_block_begin = tc_begin; g_block_end = tc_end; g_blocks_descending = true; } } g_tc_hooks;
}
extern "C" long long emu_tc_mma_count() { return ::emu::g_mma_count.load(); }
extern "C" void emu_tc_set_async_late(int v) { ::emu::g_async_late = v; }
extern "C" void emu_set_blocks_descending(int v) { ::emu::g_blocks_descending = v != 0; }
'''


def build_tc(verbose=False):
    """Host-emulated copy of csrc/spconv_tc.cu (tcgen05 / TMEM / bulk-copy kernels) over tc_emul.h."""
    os.makedirs(OUT, exist_ok=True)
    lib = os.path.join(OUT, 'libmsmd_tc_emul.so')
    deps = [os.path.join(CSRC, f) for f in ('common.cuh', 'tc_common.cuh', 'tc_trace.cuh', 'spconv_tc.cu', 'spconv_tc16.cu', 'spconv_wgrad_tc.cu', 'spconv_sb.cu')] + \
        [os.path.join(HERE, 'cuda_emul.h'), os.path.join(HERE, 'tc_emul.h'), os.path.abspath(__file__)]
    if os.path.exists(lib) and all(os.path.getmtime(lib) > os.path.getmtime(d) for d in deps):
        return lib
    _INLINED.clear()
    _INLINED.add('tc.cuh')   # replaced by tc_emul.h
    unit = translate('spconv_tc.cu') + translate('spconv_tc16.cu') + translate('spconv_wgrad_tc.cu') + \
        translate('spconv_sb.cu')   # tc_common.cuh is inlined once
    _INLINED.clear()
    # dynamic shared memory: the window tc_emul.h hands out (deliberately 16-byte aligned only)
    unit, n = re.subn(r'extern __shared__ uint8_t (\w+)\[\];', r'uint8_t* \1 = ::emu::g_dyn_smem;', unit)
    assert n >= 4
    fwd_decl = '''
extern "C" int emu_msmd_spconv_fwd_tc_ws(const float*, int, const float*, const int*, int, int, int, int, const float*,
                                         const float*, const float*, int, float*, void*, size_t, msmd_stream_t);
extern "C" size_t emu_msmd_spconv_tc_workspace(int, int);
extern "C" int emu_msmd_spconv_bwd_weight_tc_supported(int, int, int);
extern "C" size_t emu_msmd_spconv_tc16_workspace(int, int);
'''   # declared by include/msmd_b200.h in the real build
    text = PRELUDE + TC_PRELUDE + _device_helpers() + fwd_decl + unit
    cpp = os.path.join(OUT, 'emul_tc_unit.cpp')
    with open(cpp, 'w') as f:
        f.write(text)
    cmd = ['g++', '-std=c++20', '-O2', '-g', '-ffp-contract=off', '-Wno-unknown-pragmas', '-shared', '-fPIC',
           '-pthread', '-I', HERE, cpp, '-o', lib]
    if verbose:
        print(' '.join(cmd))
    subprocess.check_call(cmd)
    return lib


def build_full(verbose=False):
    """Every translation unit except points.cu (thread-block clusters) in ONE emulated image: what
    tests/tools/emu_plugin.py runs the -m gpu tests on."""
    return build_exec(verbose, name='libmsmd_full_emul.so',
                      units=['error_stub', 'voxelize.cu', 'spconv.cu', 'rulebook.cu', 'fusion.cu', 'spconv_tc.cu',
                             'spconv_tc16.cu', 'spconv_sb.cu', 'spconv_wgrad_tc.cu', 'spconv_bwd.cu', 'executor.cu',
                             'gma.cu'])


def build_exec(verbose=False, name='libmsmd_exec_emul.so', units=None):
    """Host-emulated native executor (csrc/executor.cu) with everything it calls: the bit-grid / rulebook
    kernels, the mask sort, the SIMT and tensor-core convolutions -- one library, entry points emu_msmd_*."""
    os.makedirs(OUT, exist_ok=True)
    lib = os.path.join(OUT, name)
    units = units or ['spconv.cu', 'rulebook.cu', 'fusion.cu', 'spconv_tc.cu', 'spconv_tc16.cu', 'spconv_sb.cu',
                      'executor.cu', 'gma.cu']
    stub = 'error_stub' in units
    units = [u for u in units if u != 'error_stub']
    header = os.path.join(ROOT, 'include', 'msmd_b200.h')
    deps = [os.path.join(CSRC, f) for f in units + ['common.cuh', 'tc_common.cuh', 'tc_trace.cuh', 'scan.cuh', 'sort.cuh']] + \
        [header, os.path.join(HERE, 'cuda_emul.h'), os.path.join(HERE, 'tc_emul.h'), os.path.abspath(__file__)]
    if os.path.exists(lib) and all(os.path.getmtime(lib) > os.path.getmtime(d) for d in deps):
        return lib
    _INLINED.clear()
    _INLINED.add('tc.cuh')
    body = ''
    for u in units:
        body += _rewrite_launches(_inline_headers(open(os.path.join(CSRC, u)).read()))
    _INLINED.clear()
    body, n = re.subn(r'extern __shared__ uint8_t (\w+)\[\];', r'uint8_t* \1 = ::emu::g_dyn_smem;', body)
    assert n >= 3
    if stub:   # what csrc/error.cu provides
        body = ('extern "C" int msmd_abi_version(void) { return MSMD_ABI_VERSION; }\n'
                'extern "C" unsigned long long msmd_launch_count(void) { return 0; }\n'
                'extern "C" const char* msmd_last_error(void) { return ::emu::g_error; }\n') + body
    text = '#undef MSMD_API\n' + open(header).read() + _device_helpers() + body
    # every C-ABI name (declarations of the header, definitions, calls) gets the emu_ prefix
    text = re.sub(r'(?<![\w])msmd_(\w+)\(', r'emu_msmd_\1(', text)
    text = '#define MSMD_EMUL_WITH_HEADER 1\n' + PRELUDE + TC_PRELUDE + text
    cpp = os.path.join(OUT, name.replace('lib', '').replace('.so', '') + '.cpp')
    with open(cpp, 'w') as f:
        f.write(text)
    cmd = ['g++', '-std=c++20', '-O2', '-g', '-ffp-contract=off', '-Wno-unknown-pragmas', '-shared', '-fPIC',
           '-pthread', '-I', HERE, cpp, '-o', lib]
    if verbose:
        print(' '.join(cmd))
    subprocess.check_call(cmd)
    return lib


if __name__ == '__main__':
    print(build_exec(verbose=True))
    print(build_tc(verbose=True))
    print(build(verbose=True))
