"""The one true golden vector of the reference for this path: Textures/CurlNoiseFBM.tga.
oracle restatement == shipped texture == the reference's own ImageUtils.cpp compiled verbatim (oracle/_ref)."""
import hashlib

import numpy as np
import pytest


def test_fixture_is_the_shipped_texture(assets):
    m = assets["manifest"]["CurlNoiseFBM"]
    assert hashlib.sha256(assets["curl"].tobytes()).hexdigest() == m["sha256"]
    assert m["sha256"].startswith("798ad9211f5842fe")
    assert hashlib.sha256(assets["lowres"].tobytes()).hexdigest().startswith("44448f940ff2f3ba")
    assert hashlib.sha256(assets["hires"].tobytes()).hexdigest().startswith("bd87fefa78192ef2")


def test_oracle_restatement_reproduces_shipped_texture(oracle, assets):
    got = oracle.generate_curl_noise()
    assert np.array_equal(got, assets["curl"]), f"{(got != assets['curl']).sum()} bytes differ"


def test_reference_generator_reproduces_shipped_texture(oracle, assets):
    if not oracle.have_ref_curl():
        pytest.skip("oracle/_ref/libref_curl.so not built (reference tree absent at build time)")
    got = oracle.ref_generate_curl_noise()
    assert np.array_equal(got, assets["curl"])
    assert np.array_equal(got, oracle.generate_curl_noise())


def test_hash_lattice_table_matches_reference_hash(oracle):
    """The product library evaluates the sin-hash gradient index on the host for the 26^3 lattice (capi.cu);
    this pins the oracle's hash on that lattice: indices in 0..11 and all 12 gradients in use."""
    l = oracle.lib()
    idx = np.array([[[l.om_curl_hash_index(float(x), float(y), float(z)) for x in range(-1, 25)] for y in range(-1, 25)] for z in (0, 1, 3, 6, 12)])
    assert idx.min() >= 0 and idx.max() <= 11 and len(np.unique(idx)) == 12
