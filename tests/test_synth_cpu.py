"""CPU: the synthetic-scan specification as implemented by the oracle."""
import numpy as np


def test_philox_known_answers(oracle):
    # Random123 kat_vectors, philox4x32-10
    assert oracle.philox4x32([0, 0, 0, 0], [0, 0]) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    f = 0xFFFFFFFF
    assert oracle.philox4x32([f, f, f, f], [f, f]) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert oracle.philox4x32([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]) == \
        [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_scan_is_band_consistent_and_maplike(oracle, mr):
    from maprast_b200 import synth
    seed, pal, stats, _ = synth.config_workload(2)
    full = oracle.synth_scan(seed, 0, pal, 300, 200)
    band = oracle.synth_scan(seed, 0, pal, 300, 200, line0=77, nlines=50)
    assert np.array_equal(band, full[77:127])
    other = oracle.synth_scan(seed, 1, pal, 300, 200)
    assert not np.array_equal(other, full)
    lab = oracle.categorize_u8(full, stats.mean, stats.lo, stats.hi)
    frac0 = (lab == 0).mean()
    assert 0.001 < frac0 < 0.05                 # strays (0.5 %) + box rejections
    cleaned = oracle.majority(lab, 2)
    changed = (cleaned != lab).mean()
    assert 0.01 < changed < 0.10                # ~2.5 % speckle/stray removed, boundaries adjusted
    uni = oracle.synth_scan(seed, 0, pal, 64, 64, mode=1)
    assert len(np.unique(uni.reshape(-1, 3), axis=0)) > 4000
