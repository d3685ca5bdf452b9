"""libsdr_b200 -- B200-native implementation of libsdr's receive-chain hot path
(IQBaseBand, FFT-convolution FilterNode, FM/AM/USB demodulators) behind libsdr's node interface.

The compute lives in libsdrg.so (hand-written sm_100a CUDA behind the C ABI of include/sdrg.h);
this package is the Python host-side mirror used by tests and bench.py.  There is no CPU fallback.
"""
from . import synth  # noqa: F401  (host-side synthetic workloads, numpy only)


def __getattr__(name):
    # node classes are resolved lazily so that `import libsdr_b200.synth` works without the .so
    if name in ("IQBaseBand", "BaseBand", "FMDemod", "AMDemod", "USBDemod", "RxChain", "FFTPlan", "FilterNode", "ChannelBank", "Config", "ConfigError"):
        from . import nodes
        return getattr(nodes, name)
    raise AttributeError(name)
