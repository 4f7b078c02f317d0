"""Network description: blob topology of the reference's ``models/vgg19.prototxt`` (lines 3-337),
a small prototxt reader to validate a user-supplied file against it, synthetic weights for
offline runs, and a dependency-free ``.caffemodel`` (protobuf wire format) reader/writer.
"""
from collections import OrderedDict
import math
import re
import struct

import numpy as np

# (blob name, kind, channels); kinds: 'input', 'conv' (3x3 pad 1 + in-place ReLU), 'pool' (2x2/2 MAX)
TOPOLOGY = [('data', 'input', 3)]
for _b, (_n, _c) in enumerate(((2, 64), (2, 128), (4, 256), (4, 512), (4, 512)), 1):
    TOPOLOGY += [('conv%d_%d' % (_b, _i), 'conv', _c) for _i in range(1, _n + 1)]
    TOPOLOGY.append(('pool%d' % _b, 'pool', _c))
BLOBS = [t[0] for t in TOPOLOGY]
BLOB_INDEX = {n: i for i, n in enumerate(BLOBS)}
CONVS = [t[0] for t in TOPOLOGY if t[1] == 'conv']


def conv_shapes():
    """conv name -> (cout, cin)."""
    out, cin = OrderedDict(), 3
    for name, kind, c in TOPOLOGY:
        if kind == 'conv':
            out[name] = (c, cin)
        cin = c
    return out


def synthetic_weights(seed=0):
    """He-normal stand-in for ``vgg19.caffemodel`` (not downloadable offline): one RandomState,
    convs in prototxt order, W (Cout, Cin, 3, 3) ~ N(0, 2/(9 Cin)), b ~ 0.1 N(0, 1), fp32."""
    rs = np.random.RandomState(seed)
    params = OrderedDict()
    for name, (cout, cin) in conv_shapes().items():
        w = (rs.randn(cout, cin, 3, 3) * math.sqrt(2.0 / (9 * cin))).astype(np.float32)
        b = (rs.randn(cout) * 0.1).astype(np.float32)
        params[name] = (w, b)
    return params


# ------------------------------------------------------------------------------- prototxt
def parse_prototxt_layers(text):
    """Minimal reader for the subset of prototxt the reference model uses: returns a list of dicts
    with name, type, bottom, top and (conv) num_output/pad/kernel_size, (pool) pool/kernel_size/stride."""
    layers = []
    for body in _layer_bodies(text):
        d = {'_body': body}
        for key in ('name', 'type', 'bottom', 'top'):
            m = re.search(r'\b%s\s*:\s*"([^"]*)"' % key, body)
            if m:
                d[key] = m.group(1)
        for key in ('num_output', 'pad', 'kernel_size', 'stride'):
            m = re.search(r'\b%s\s*:\s*(\d+)' % key, body)
            if m:
                d[key] = int(m.group(1))
        m = re.search(r'\bpool\s*:\s*(\w+)', body)
        if m:
            d['pool'] = m.group(1)
        layers.append(d)
    return layers


def _layer_bodies(text):
    i = 0
    while True:
        m = re.compile(r'\blayers?\s*\{').search(text, i)
        if not m:
            return
        depth, j = 1, m.end()
        while depth and j < len(text):
            depth += {'{': 1, '}': -1}.get(text[j], 0)
            j += 1
        yield text[m.end():j - 1]
        i = j


_V1_TYPES = {'CONVOLUTION': 'Convolution', 'RELU': 'ReLU', 'POOLING': 'Pooling', 'INNER_PRODUCT': 'InnerProduct',
             'DROPOUT': 'Dropout', 'SOFTMAX': 'Softmax', 'DATA': 'Data'}


def _top_level_inputs(text):
    """``input: "data"`` declarations outside any layer block (deploy-style prototxts, both dialects)."""
    depth, out, i = 0, [], 0
    for m in re.finditer(r'[{}]|\binput\s*:\s*"([^"]*)"', text):
        tok = m.group(0)
        if tok == '{':
            depth += 1
        elif tok == '}':
            depth -= 1
        elif depth == 0:
            out.append(m.group(1))
    return out


def net_from_prototxt(text, strict=True):
    """The blob list a prototxt describes, as far as this engine implements it.

    Accepts the current dialect (``layer { type: "Convolution" }``, an ``Input`` layer) and the legacy V1 one
    (``layers { type: CONVOLUTION }``, top-level ``input: "data"``), i.e. the reference's ``models/vgg19.prototxt``
    (vgg19.prototxt:1-337) as well as the model zoo's VGG-19 deploy files.  The convolutional stack must be a PREFIX
    of the VGG-19 topology above (3x3 pad-1 stride-1 convolutions with in-place ReLUs, 2x2/2 MAX pools, the same blob
    names, order and widths) -- a network cut earlier (say after conv4_2) is fine and simply exposes fewer blobs.
    ``strict=False`` additionally tolerates a tail the style-transfer path never evaluates (InnerProduct / Dropout /
    Softmax after the last pool: the reference's file is the zoo's with exactly that tail removed) and returns its
    layer names.  Returns ``(blobs, ignored)``; raises ValueError for anything else."""
    blobs = [(name, 'input', 3) for name in _top_level_inputs(text)[:1]]
    ignored = []
    for layer in parse_prototxt_layers(text):
        kind = layer.get('type')
        if kind is None:
            m = re.search(r'\btype\s*:\s*([A-Z_]+)', layer.get('_body', ''))
            kind = m.group(1) if m else None
        kind = _V1_TYPES.get(kind, kind)
        if ignored:                                  # once the tail has begun nothing of it is interpreted
            ignored.append(layer.get('name', '?'))
            continue
        if kind == 'Input':
            blobs.append((layer['top'], 'input', 3))
        elif kind == 'Convolution':
            if layer.get('kernel_size') != 3 or layer.get('pad') != 1 or layer.get('stride', 1) != 1:
                raise ValueError('unsupported convolution geometry in %s' % layer.get('name'))
            blobs.append((layer['top'], 'conv', layer['num_output']))
        elif kind == 'Pooling':
            if layer.get('pool') != 'MAX' or layer.get('kernel_size') != 2 or layer.get('stride') != 2:
                raise ValueError('unsupported pooling in %s' % layer.get('name'))
            blobs.append((layer['top'], 'pool', blobs[-1][2]))
        elif kind == 'ReLU':
            if layer.get('bottom') != layer.get('top') or not blobs or layer.get('top') != blobs[-1][0]:
                raise ValueError('ReLU %s is not in place on the preceding conv' % layer.get('name'))
        elif not strict and kind in ('InnerProduct', 'Dropout', 'Softmax') and blobs and blobs[-1][1] == 'pool':
            ignored.append(layer.get('name', '?'))
        else:
            raise ValueError('unsupported layer type %r' % kind)
    if len(blobs) < 2 or blobs != TOPOLOGY[:len(blobs)]:
        raise ValueError('prototxt does not describe (a prefix of) the truncated VGG-19 topology')
    return blobs, ignored


def check_prototxt(text):
    """Raise ValueError unless the prototxt describes exactly the truncated VGG-19 of the reference
    (same blob names and order, 3x3 pad-1 convs, 2x2/2 MAX pools, in-place ReLUs)."""
    blobs, _ = net_from_prototxt(text, strict=True)
    if blobs != TOPOLOGY:
        raise ValueError('prototxt does not describe the truncated VGG-19 topology')
    return True


# ------------------------------------------------------------------------------- caffemodel
def _varint(buf, pos):
    shift = val = 0
    while True:
        b = buf[pos]
        pos += 1
        val |= (b & 0x7F) << shift
        if not b & 0x80:
            return val, pos
        shift += 7


def _fields(buf):
    """Yield (field number, wire type, value) of one protobuf message; value is int or memoryview."""
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            val, pos = _varint(buf, pos)
        elif wt == 1:
            val, pos = buf[pos:pos + 8], pos + 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            val, pos = buf[pos:pos + ln], pos + ln
        elif wt == 5:
            val, pos = buf[pos:pos + 4], pos + 4
        else:
            raise ValueError('unsupported protobuf wire type %d' % wt)
        yield fno, wt, val


def _blob(buf):
    """BlobProto: shape = 7 (BlobShape.dim = 1), data = 5 (packed float), legacy num/channels/height/width = 1-4."""
    dims, legacy, chunks = [], {}, []
    for fno, wt, val in _fields(buf):
        if fno == 7 and wt == 2:
            for f2, w2, v2 in _fields(val):
                if f2 == 1 and w2 == 2:
                    p = 0
                    while p < len(v2):
                        d, p = _varint(v2, p)
                        dims.append(d)
                elif f2 == 1 and w2 == 0:
                    dims.append(v2)
        elif fno == 5 and wt == 2:
            chunks.append(np.frombuffer(val, dtype='<f4'))
        elif fno == 5 and wt == 5:
            chunks.append(np.frombuffer(val, dtype='<f4'))
        elif fno in (1, 2, 3, 4) and wt == 0:
            legacy[fno] = val
    data = np.concatenate(chunks) if chunks else np.zeros(0, np.float32)
    if not dims and legacy:
        dims = [legacy.get(k, 1) for k in (1, 2, 3, 4)]
    return data.reshape(dims) if dims else data


def read_caffemodel(path):
    """conv name -> (W (Cout, Cin, 3, 3), b (Cout)) from a NetParameter file: ``layer`` = field 100
    (LayerParameter: name 1, type 2, blobs 7) or legacy V1 ``layers`` = field 2 (name 4, blobs 6)."""
    with open(path, 'rb') as f:
        buf = memoryview(f.read())
    params = OrderedDict()
    for fno, wt, val in _fields(buf):
        if wt != 2 or fno not in (100, 2):
            continue
        name_f, blobs_f = (1, 7) if fno == 100 else (4, 6)
        name, blobs = None, []
        for f2, w2, v2 in _fields(val):
            if f2 == name_f and w2 == 2:
                name = bytes(v2).decode()
            elif f2 == blobs_f and w2 == 2:
                blobs.append(_blob(v2))
        if name in CONVS and len(blobs) >= 1:
            w = np.ascontiguousarray(blobs[0], np.float32)
            b = np.ascontiguousarray(blobs[1], np.float32).reshape(-1) if len(blobs) > 1 else \
                np.zeros(w.shape[0], np.float32)
            params[name] = (w.reshape(w.shape[-4:]) if w.ndim >= 4 else w, b)
    shapes = conv_shapes()
    for name, (cout, cin) in shapes.items():
        if name not in params:
            raise ValueError('caffemodel has no weights for %s' % name)
        if params[name][0].shape != (cout, cin, 3, 3) or params[name][1].shape != (cout,):
            raise ValueError('%s: unexpected blob shapes %s / %s' % (name, params[name][0].shape,
                                                                    params[name][1].shape))
    return OrderedDict((n, params[n]) for n in shapes)


def _enc_varint(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        out.append(b | (0x80 if v else 0))
        if not v:
            return bytes(out)


def _enc_field(fno, payload):
    return _enc_varint((fno << 3) | 2) + _enc_varint(len(payload)) + payload


def write_caffemodel(path, params, net_name='vgg19_truncated'):
    """Inverse of ``read_caffemodel`` (modern ``layer`` = 100 form); used by tests and to hand
    synthetic weights to other Caffe-format consumers (e.g. ``cv2.dnn.readNetFromCaffe``)."""
    out = bytearray(_enc_field(1, net_name.encode()))
    for name, (w, b) in params.items():
        layer = bytearray(_enc_field(1, name.encode()) + _enc_field(2, b'Convolution'))
        for arr in (w, b):
            arr = np.ascontiguousarray(arr, '<f4')
            shape = _enc_field(1, b''.join(_enc_varint(int(d)) for d in arr.shape))
            layer += _enc_field(7, _enc_field(7, shape) + _enc_field(5, arr.tobytes()))
        out += _enc_field(100, bytes(layer))
    with open(path, 'wb') as f:
        f.write(bytes(out))
