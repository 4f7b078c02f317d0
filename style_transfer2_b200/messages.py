"""Wire protocol of the worker: the pickleable message types the reference exchanges over PyZMQ
(``messages.py:13-172``).  Pickles carry the module path ``messages.<Class>``, so the drop-in worker
installs this module under the top-level name ``messages`` (see ``install_as_toplevel``) before it
opens its sockets; the unmodified ``app.py`` then talks to it without noticing.

Same class names, constructor arguments and attributes as upstream; ``SetOptimizer.classes`` points
at this package's device-resident optimizers.
"""
import inspect
import logging
import sys

import numpy as np

from . import optimizers

logger = logging.getLogger('messages')


def _short(value):
    if isinstance(value, np.ndarray):
        return '<ndarray, shape: %s, dtype: %s>' % (value.shape, value.dtype)
    return repr(value)


class Message:
    """Base of everything sent with ``send_pyobj`` (messages.py:13-35)."""
    debug = False

    def __repr__(self):
        fields = ', '.join('%s=%s' % (k, _short(v)) for k, v in sorted(vars(self).items()))
        return '%s(%s)' % (type(self).__name__, fields)

    def _debug(self):
        if not self.debug:
            return
        frame = inspect.currentframe()
        try:
            site = frame.f_back.f_back
            logger.debug('%s created on line %d of %s: %r', type(self).__name__, site.f_lineno,
                         site.f_code.co_filename, self)
        finally:
            del frame


def _plain(name, doc, *fields, **defaults):
    """Message class whose constructor stores its arguments as same-named attributes."""
    def __init__(self, *args, **kwargs):
        values = dict(defaults)
        if len(args) > len(fields):
            raise TypeError('%s takes at most %d arguments' % (name, len(fields)))
        values.update(zip(fields, args))
        for k, v in kwargs.items():
            if k not in fields:
                raise TypeError('%s got an unexpected argument %r' % (name, k))
            values[k] = v
        missing = [f for f in fields if f not in values]
        if missing:
            raise TypeError('%s missing arguments: %s' % (name, ', '.join(missing)))
        for f in fields:
            setattr(self, f, values[f])
        self._debug()
    return type(name, (Message,), {'__init__': __init__, '__doc__': doc, '__module__': __name__})


# router-facing (messages.py:38-53, 82-86) -- kept so that a shared ``messages`` module stays complete
AppDown = _plain('AppDown', 'App -> router: the app is shutting down.', 'addr', 'app_id')
AppUp = _plain('AppUp', 'App -> router: the app is up.', 'addr', 'host', 'port', 'app_id')
Reset = _plain('Reset', 'Router -> app: reset your state and your worker.')

# worker-facing
GetImages = _plain('GetImages', 'Worker -> app: image slots are missing, send them (messages.py:56-61).')
Iterate = _plain('Iterate', 'Worker -> app: a new iterate -- float32-convertible HxWx3 RGB image, iterate '
                 'count since reset, trace dict (messages.py:64-74).', 'image', 'i', 'trace')
PauseIteration = _plain('PauseIteration', 'App -> worker: pause (messages.py:77-80).')
Shutdown = _plain('Shutdown', 'Either way: shut down (messages.py:152-155).')
StartIteration = _plain('StartIteration', 'App -> worker: start iterating (messages.py:159-162).')


class SetImages(Message):
    """App -> worker: fill image slots (messages.py:89-110).  Images are HxWx3 RGB arrays; ``None``
    leaves a slot alone; ``SetImages.RESAMPLE`` asks the worker to resample its own copy to
    ``size``; ``reset_state`` clears the optimizer and the iterate counter."""
    RESAMPLE = 1

    def __init__(self, size=None, input_image=None, content_image=None, style_image=None, reset_state=False):
        self.size = size
        self.input_image = input_image
        self.content_image = content_image
        self.style_image = style_image
        self.reset_state = reset_state
        self._debug()


class SetOptimizer(Message):
    """App -> worker: optimizer type and step size (messages.py:113-128)."""
    classes = {'adam': optimizers.AdamOptimizer, 'lbfgs': optimizers.LBFGSOptimizer}
    step_sizes = {'adam': 10, 'lbfgs': 1}

    def __init__(self, optimizer, step_size=None):
        if optimizer not in self.classes:
            raise ValueError('Invalid optimizer type')
        self.optimizer = optimizer
        self.step_size = step_size if step_size else self.step_sizes[optimizer]
        self._debug()


class SetWeights(Message):
    """App -> worker: ``weights[loss][layer]`` table and the scalar pixel-space parameters
    (messages.py:131-149)."""
    loss_names = ('content', 'style', 'deepdream')
    scalar_loss_names = ('tv', 'tv_power', 'p', 'p_power')

    def __init__(self, weights, params):
        self.weights = weights
        self.params = params
        self._debug()


class WorkerReady(Message):
    """Worker -> app: ready; carries the layer names (messages.py:166-172)."""

    def __init__(self, layers=None):
        self.layers = list(layers) if layers is not None else []
        self._debug()


WORKER_FACING = (GetImages, Iterate, PauseIteration, SetImages, SetOptimizer, SetWeights, Shutdown,
                 StartIteration, WorkerReady)


def install_as_toplevel():
    """Make ``import messages`` / ``pickle`` resolve to this module, and re-home the classes so
    that what we pickle is readable by a process that only has the reference's ``messages.py``."""
    mod = sys.modules[__name__]
    sys.modules['messages'] = mod
    for obj in list(vars(mod).values()):
        if isinstance(obj, type) and issubclass(obj, Message):
            obj.__module__ = 'messages'
    return mod
