"""ofdm_uhd_b200 -- B200-native (sm_100a) OFDM baseband hot path with the Python surface of
rubiruchi/ofdm_uhd (ofdm.py / transmit_path.py / receive_path.py / ofdm_packet_utils.py / psk.py / qam.py).

The modules can be imported as ``ofdm_uhd_b200.ofdm`` or, like the reference's flat script directory,
by putting this directory on ``sys.path`` and writing ``import ofdm``.
"""
__version__ = "0.1.0"
