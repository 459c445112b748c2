"""Host marker walk (JpegDecoder.Identify / marker loop mirror) -- CPU only.
Known answers: tests/JpegLibrary.Tests/Decoder/MetadataIdentifyTests.cs:19-130."""
import pytest

import jpeglibrary_b200 as J
import oracle_ffi as O
import synth
from conftest import golden_bytes


@pytest.mark.parametrize("name", ["cramps.jpg", "testorig12.jpg", "progress.jpg", "yellowcat_progressive_restart.jpg"])
def test_identify_known_answers(name, golden):
    k = golden["identify"][name]
    dec = J.JpegDecoder()
    dec.SetInput(golden_bytes(name))
    consumed = dec.Identify()
    assert (dec.Width, dec.Height, dec.NumberOfComponents, dec.Precision) == (k["Width"], k["Height"], k["NumberOfComponents"], k["Precision"])
    assert consumed == k["JpegStreamSize"]


def test_identify_requires_input():
    dec = J.JpegDecoder()
    with pytest.raises(J.InvalidOperationException):
        dec.Identify()
    with pytest.raises(J.InvalidOperationException):
        dec.Width
    dec.SetInput(b"\x00\x01\x02\x03")
    with pytest.raises(J.InvalidDataException):
        dec.Identify()


@pytest.mark.parametrize("name", ["lake.jpg", "cramps.jpg", "testorig12.jpg", "progress.jpg", "yellowcat_progressive_restart.jpg"])
def test_descriptor_matches_oracle_headers(name):
    blob = golden_bytes(name)
    p = J.Parsed(blob)
    d = p.desc
    o = O.decode(blob, want_rgb=False)
    assert (d.width, d.height, d.component_count, d.precision, d.sof) == (o.width, o.height, o.ncomp, o.precision, o.sof)
    assert d.scan_count == o.nscans
    for c in range(o.ncomp):
        assert (d.h[c], d.v[c]) == (o.comp_h[c], o.comp_v[c])
        assert list(d.quant[c]) == list(o.qt[c])
    for i in range(o.nscans):
        s, t = d.scans[i], o.scans[i]
        assert (s.ss, s.se, s.ah, s.al, s.restart_interval, s.entropy_offset) == (t.ss, t.se, t.ah, t.al, t.restart_interval, t.entropy_offset)
        assert blob[s.entropy_offset + s.entropy_length] == 0xFF  # the scan ends at a marker


def test_restart_interval_and_sampling_getters():
    blob = synth.synth_jpeg(2, 80, 48, restart_rows=1)
    dec = J.JpegDecoder()
    dec.SetInput(blob)
    dec.Identify()
    assert dec.GetMaximumHorizontalSampling() == 2 and dec.GetMaximumVerticalSampling() == 2
    assert dec.GetHorizontalSampling(1) == 1 and dec.GetVerticalSampling(0) == 2
    with pytest.raises(J.ArgumentException):
        dec.GetHorizontalSampling(3)
    assert J.Parsed(blob).desc.scans[0].restart_interval == 5  # one MCU row of an 80-px-wide 4:2:0 image


def test_decode_call_order_errors():
    dec = J.JpegDecoder()
    with pytest.raises(J.InvalidOperationException):
        dec.Decode()
    dec.SetInput(synth.synth_jpeg(2, 32, 32))
    with pytest.raises(J.InvalidOperationException):
        dec.Decode()  # "The output buffer is not specified."
    with pytest.raises(J.ArgumentException):
        dec.SetOutputWriter(None)
