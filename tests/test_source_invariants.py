"""Static checks of invariants the CUDA sources rely on (no GPU, no compiler)."""
import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, 'style_transfer2_b200', 'csrc')


def _sources():
    return {os.path.basename(p): open(p).read() for p in sorted(glob.glob(os.path.join(CSRC, '*.cu')))}


def _kernel_bodies(text):
    """name -> source text of every __global__ function (up to the next __global__ or the end of the file)."""
    out = {}
    parts = re.split(r'(?=__global__)', text)
    for part in parts[1:]:
        head = re.sub(r'__(launch_bounds|cluster_dims)__\s*\([^)]*\)', ' ', part[:600])
        # the kernel name is the identifier right before the parameter list that follows `void`
        m = re.search(r'\bvoid\b\s+([A-Za-z_][A-Za-z0-9_]*)\s*\(', head)
        if m:
            out[m.group(1)] = part
    return out


def test_every_kernel_launched_with_programmatic_serialisation_waits_for_its_predecessor():
    """st2_launch_pdl(ctx, true, K, ...) lets K start while its predecessor drains; K must execute pdl_wait() before
    it touches global memory, and the chain of kernels is only transitive because EVERY link waits (st2_common.cuh)."""
    launched, bodies = set(), {}
    for name, text in _sources().items():
        bodies.update(_kernel_bodies(text))
        for m in re.finditer(r'st2_launch_pdl\(\s*ctx\s*,\s*true\s*,\s*([A-Za-z_][A-Za-z0-9_]*)', text):
            launched.add(m.group(1))
        # the L-BFGS passes go through a macro that forwards the kernel name
        for m in re.finditer(r'LAUNCH_V\(\s*([A-Za-z_][A-Za-z0-9_]*)', text):
            if 'st2_launch_pdl(ctx, true, kernel<' in text:
                launched.add(m.group(1))
    launched.discard('kernel')
    assert len(launched) >= 20, sorted(launched)
    missing = [k for k in sorted(launched) if k not in bodies]
    assert not missing, 'kernels not found: %s' % missing
    for k in sorted(launched):
        body = bodies[k]
        assert 'pdl_wait()' in body, '%s is launched with programmatic serialisation but never calls pdl_wait()' % k
        assert 'pdl_trigger()' in body, '%s never lets its successor in (pdl_trigger)' % k
        # nothing that reads global memory may come before the wait: the only loads allowed above it are of kernel
        # parameters; __ldg / ld.global / tma loads before pdl_wait() would race with the predecessor
        head = body[:body.index('pdl_wait()')]
        for needle in ('__ldg(', 'tma_load', 'atomicAdd(', 'cp.async.bulk'):
            assert needle not in head or k.startswith('tc_conv') and needle == 'atomicAdd(' and 'halo_push_prologue' in head, \
                '%s touches global memory (%s) before pdl_wait()' % (k, needle)


def test_halo_pushing_launches_stay_plain():
    """Kernels that push halo rows in their prologue read their input before the setup barrier: they must be launched
    WITHOUT programmatic serialisation (st2_conv_tc.cu: the <..., true> instantiations)."""
    text = _sources()['st2_conv_tc.cu']
    for m in re.finditer(r'st2_launch_pdl\(\s*ctx\s*,\s*(true|false)\s*,\s*(tc_conv[A-Za-z0-9_]*)<([^>]*)>', text):
        pdl, kernel, targs = m.group(1), m.group(2), [a.strip() for a in m.group(3).split(',')]
        halo = targs[1] if kernel == 'tc_conv2_kernel' else targs[-1]
        assert (halo == 'true') == (pdl == 'false'), m.group(0)
