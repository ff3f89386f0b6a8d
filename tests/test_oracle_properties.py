"""Property tests of the oracle (CPU): every stream the restatement emits is a valid bzip2 stream for libbz2, the
decode authority, and its own decoder agrees.  Inputs the reference itself rejects (input ending in a run
of exactly four) raise RefPanic and are skipped."""
import bz2

from hypothesis import HealthCheck, assume, given, settings, strategies as st


@st.composite
def run_heavy(draw):
    """Byte strings made of runs (lengths around the RLE1 thresholds 3/4/5 and 255/256) over a small alphabet."""
    pieces = draw(st.lists(st.tuples(st.integers(0, 5), st.sampled_from([1, 2, 3, 4, 5, 6, 7, 50, 254, 255, 256, 259, 600])),
                           min_size=0, max_size=40))
    return b"".join(bytes([97 + v]) * n for v, n in pieces)


@settings(max_examples=120, deadline=None, derandomize=True, suppress_health_check=[HealthCheck.too_slow])
@given(data=st.one_of(st.binary(max_size=3000), run_heavy()), level=st.sampled_from([1, 9]))
def test_oracle_streams_decode_with_libbz2(ref, data, level):
    # an input that ends in four equal bytes makes the reference panic or emit an invalid stream (the
    # restatement follows it there; the engine deliberately does not)
    assume(not (len(data) >= 4 and data[-4:] == data[-1:] * 4))
    try:
        stream = ref.compress_stream(data, level, ref.SPEC_FAST)
    except ref.RefPanic:
        return
    assert bz2.decompress(stream) == data
    assert ref.decompress_stream(stream) == data


@settings(max_examples=60, deadline=None, derandomize=True, suppress_health_check=[HealthCheck.too_slow])
@given(data=st.binary(min_size=1, max_size=2000))
def test_oracle_bwt_modes_agree(ref, data):
    k0, b0, _ = ref.bwt_encode(data, ref.SPEC)
    k1, b1, _ = ref.bwt_encode(data, ref.SPEC_FAST)
    assert (k0, b0) == (k1, b1)
    assert ref.bwt_decode(k0, b0) == data
