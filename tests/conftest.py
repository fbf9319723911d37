import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure).  Built on demand; oracle/_ref only when the
    reference tree is present (never on the GPU box)."""
    from oracle import pyoracle

    pyoracle.build(ref=True)
    return pyoracle


@pytest.fixture(scope="session")
def rx_params():
    from gr4_packet_modem_b200.firdes import BPSK, SYNCWORD, unit_energy_rrc

    return dict(rrc_taps=unit_energy_rrc(), syncword=SYNCWORD, constellation=BPSK)
