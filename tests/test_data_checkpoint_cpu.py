"""CPU tests of the data-format and checkpoint rows either side of the hot path."""
import gzip
import struct

import numpy as np
import pytest


def test_tokenizer_matches_reference_contract():
    from progen_b200.data import encode_tokens, decode_tokens, collate
    assert encode_tokens('[Tax=Mammalia] #')[:4] == [ord(c) + 1 for c in '[Tax']
    assert decode_tokens(np.array(encode_tokens('MKV'))) == 'MKV'
    rows = collate([b'ABC', b'ABCDEFGH'], seq_len=5)
    assert rows.dtype == np.uint16 and rows.shape == (2, 6)
    np.testing.assert_array_equal(rows[0], [0, 66, 67, 68, 0, 0])          # BOS, bytes + 1, zero padding
    np.testing.assert_array_equal(rows[1], [0, 66, 67, 68, 69, 70])        # truncated to seq_len


def test_crc32c_known_answers():
    from progen_b200.data import _crc32c
    assert _crc32c(b'') == 0
    assert _crc32c(b'123456789') == 0xE3069283                              # CRC-32C check value
    assert _crc32c(bytes(32)) == 0x8A9136AA                                 # RFC 3720 B.4: 32 zero bytes


def test_tfrecord_round_trip_and_format(tmp_path):
    from progen_b200.data import write_tfrecords, read_tfrecords, iterator_from_tfrecords_folder, encode_example, decode_example
    seqs = [b'[tax=Mammalia] # MKVLAAGIVGL', b'# AAAA [tax=Bacteria]', b'X' * 300]
    path = tmp_path / '0.3.train.tfrecord.gz'
    write_tfrecords(str(path), seqs)
    assert list(read_tfrecords(str(path))) == seqs
    # wire format: first record header is the little-endian payload length
    raw = gzip.open(path, 'rb').read()
    ln, = struct.unpack('<Q', raw[:8])
    assert decode_example(raw[12:12 + ln]) == seqs[0] and encode_example(seqs[0]) == raw[12:12 + ln]
    # protobuf payload layout of tf.train.Example with one bytes feature 'seq'
    assert raw[12] == 0x0A and b'seq' in raw[12:12 + ln]
    write_tfrecords(str(tmp_path / '1.2.valid.tfrecord.gz'), seqs[:2])
    n, it = iterator_from_tfrecords_folder(str(tmp_path), 'train')
    assert n == 3
    batches = list(it(seq_len=16, batch_size=2))
    assert [b.shape for b in batches] == [(2, 17), (1, 17)]
    assert batches[0][0, 0] == 0 and batches[0][0, 1] == ord('[') + 1
    assert len(list(it(seq_len=16, batch_size=2, skip=2))) == 1
    nv, itv = iterator_from_tfrecords_folder(str(tmp_path), 'valid')
    g = itv(seq_len=8, batch_size=2, loop=True)
    assert nv == 2 and next(g).shape == (2, 9) and next(g).shape == (2, 9)
    # corruption is detected
    bad = bytearray(raw)
    bad[20] ^= 0xFF
    with gzip.open(tmp_path / 'bad.1.train.tfrecord.gz', 'wb') as f:
        f.write(bytes(bad))
    with pytest.raises(ValueError):
        list(read_tfrecords(str(tmp_path / 'bad.1.train.tfrecord.gz')))


def test_checkpoint_package_round_trip_and_keep_n(tmp_path):
    from progen_b200.checkpoint import get_checkpoint_fns
    reset, get_last, save = get_checkpoint_fns(str(tmp_path / 'ckpts'))
    assert get_last() is None
    params = {'pro_gen_base/~/embed': {'embeddings': np.arange(12, dtype=np.float32).reshape(3, 4)}}
    for i in range(4):
        save({'next_seq_index': 16 * (i + 1), 'params': params, 'optim_state': {'count': i}, 'model_config': {'dim': 4},
              'run_id': None}, keep_last_n=2)
    files = sorted((tmp_path / 'ckpts').glob('ckpt_*.pkl'))
    assert 2 <= len(files) <= 3                       # the reference prunes BEFORE counting the new file (checkpoint.py:36-37)
    last = get_last()
    assert last['next_seq_index'] == 64 and last['optim_state']['count'] == 3 and last['model_config'] == {'dim': 4}
    np.testing.assert_array_equal(last['params']['pro_gen_base/~/embed']['embeddings'], params['pro_gen_base/~/embed']['embeddings'])
    reset()
    assert get_last() is None
    with pytest.raises(NotImplementedError):
        get_checkpoint_fns('gs://bucket')


def test_param_tree_matches_reference_module_paths():
    """The product's own parameter tree (names, shapes) equals the one the reference source builds (golden cases)."""
    from progen_b200 import ProGen
    from golden_util import CASES, load_case
    for name in ('tiny_glu_sgu', 'tiny_gelu_sgu'):
        cfg, params, data, g = load_case(name)
        shapes = ProGen(**CASES[name]).param_shapes()
        assert {m: {k: tuple(v.shape) for k, v in d.items()} for m, d in params.items()} == shapes
        init = ProGen(**CASES[name]).init(0)
        assert {m: {k: v.shape for k, v in d.items()} for m, d in init.items()} == shapes
        sw = [v['spatial_weights'] for m, v in init.items() if 'spatial_weights' in v][0]
        assert np.abs(sw).max() <= 1e-3 / cfg['seq_len'] + 1e-12          # U(+-eps/n), progen.py:172-176
