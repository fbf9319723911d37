"""gr4_packet_modem_b200 — B200 (sm_100a) receiver-synchronisation hot path of gr4-packet-modem.

The product is csrc/ -> libb200sync.so (hand-written CUDA behind the C ABI of
include/b200sync.h).  This package is the thin host-side mirror of the reference's block
interface on top of that ABI; there is no CPU fallback — importing the blocks without the
built library raises.
"""
from .firdes import root_raised_cosine  # noqa: F401
from .blocks import (SyncwordDetection, SyncwordDetectionMulti, DetectionRecord, SyncwordTag, FrontEnd, PfbArbResampler,  # noqa: F401
                     Rotator, SymbolFilter, SyncwordDetectionFilter, CoarseFrequencyCorrection, SyncwordWipeoff, CostasLoop,
                     host_register, host_unregister)
