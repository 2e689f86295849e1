"""TFRecord / `tf.train.Example` codec without TensorFlow (transformer/tfrecord.py) and the
reference's record layout (`BatchUniversalTransformer.encode / decode_protobuf`,
/root/reference/tensoralloy/transformer/universal.py:1177-1330, base.py:365-437).

Pins: CRC-32C's published check value and RFC 3720 B.4 vectors; the framing of a record
written by TensorFlow (bytes stated below); the message codec in both directions against
google.protobuf driven by the example.proto / feature.proto schema built from descriptors."""
import struct

import numpy as np
import pytest

from oracle import neighbor as onl
from tensoralloy_b200.atoms import Atoms, bulk_fcc
from tensoralloy_b200.io.read import Dataset
from tensoralloy_b200.precision import precision_scope
from tensoralloy_b200.transformer import BatchUniversalTransformer
from tensoralloy_b200.transformer import tfrecord as tfr


def test_crc32c_known_answers():
    assert tfr.crc32c(b'123456789') == 0xE3069283
    # RFC 3720 B.4
    assert tfr.crc32c(bytes(32)) == 0x8A9136AA
    assert tfr.crc32c(bytes([0xFF] * 32)) == 0x62A8AB43
    assert tfr.crc32c(bytes(range(32))) == 0x46DD794E
    assert tfr.crc32c(bytes(range(31, -1, -1))) == 0x113FDB5C
    # slicing-by-8 body == bytewise tail, any length / split
    rng = np.random.default_rng(3)
    data = rng.integers(0, 256, 1021, dtype=np.uint8).tobytes()
    t0 = tfr._crc_table()[0]
    c = 0xFFFFFFFF
    for b in data:
        c = t0[(c ^ b) & 0xFF] ^ (c >> 8)
    assert tfr.crc32c(data) == c ^ 0xFFFFFFFF
    assert tfr.crc32c(b'') == 0


def test_record_framing(tmp_path):
    # masked crc of the 8-byte length header of a 3-byte record, and the whole frame
    path = str(tmp_path / 'a.tfrecords')
    recs = [b'abc', b'', bytes(range(256)) * 5]
    assert tfr.write_tfrecords(path, recs) == 3
    raw = open(path, 'rb').read()
    assert len(raw) == sum(16 + len(r) for r in recs)
    assert struct.unpack_from('<Q', raw, 0)[0] == 3
    head_crc = struct.unpack_from('<I', raw, 8)[0]
    c = tfr.crc32c(struct.pack('<Q', 3))
    assert head_crc == (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF
    assert raw[12:15] == b'abc'
    assert list(tfr.read_tfrecords(path)) == recs
    # corruption is detected
    bad = bytearray(raw)
    bad[13] ^= 1
    open(path, 'wb').write(bytes(bad))
    with pytest.raises(IOError):
        list(tfr.read_tfrecords(path))
    open(path, 'wb').write(raw[:-3])
    with pytest.raises(IOError):
        list(tfr.read_tfrecords(path))


def _example_schema():
    """tensorflow/core/example/{feature,example}.proto as dynamic messages."""
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    F = descriptor_pb2.FieldDescriptorProto
    fd = descriptor_pb2.FileDescriptorProto(name='tab_example.proto', package='tabtest',
                                            syntax='proto3')

    def msg(name):
        m = fd.message_type.add()
        m.name = name
        return m

    def field(m, name, number, ftype, label=F.LABEL_OPTIONAL, type_name=None, packed=None,
              oneof=None):
        f = m.field.add()
        f.name, f.number, f.type, f.label = name, number, ftype, label
        if type_name:
            f.type_name = type_name
        if packed is not None:
            f.options.packed = packed
        if oneof is not None:
            f.oneof_index = oneof
        return f

    field(msg('BytesList'), 'value', 1, F.TYPE_BYTES, F.LABEL_REPEATED)
    field(msg('FloatList'), 'value', 1, F.TYPE_FLOAT, F.LABEL_REPEATED, packed=True)
    field(msg('Int64List'), 'value', 1, F.TYPE_INT64, F.LABEL_REPEATED, packed=True)
    feat = msg('Feature')
    feat.oneof_decl.add().name = 'kind'
    field(feat, 'bytes_list', 1, F.TYPE_MESSAGE, type_name='.tabtest.BytesList', oneof=0)
    field(feat, 'float_list', 2, F.TYPE_MESSAGE, type_name='.tabtest.FloatList', oneof=0)
    field(feat, 'int64_list', 3, F.TYPE_MESSAGE, type_name='.tabtest.Int64List', oneof=0)
    feats = msg('Features')
    entry = feats.nested_type.add()
    entry.name = 'FeatureEntry'
    entry.options.map_entry = True
    field(entry, 'key', 1, F.TYPE_STRING)
    field(entry, 'value', 2, F.TYPE_MESSAGE, type_name='.tabtest.Feature')
    field(feats, 'feature', 1, F.TYPE_MESSAGE, F.LABEL_REPEATED,
          type_name='.tabtest.Features.FeatureEntry')
    ex = msg('Example')
    field(ex, 'features', 1, F.TYPE_MESSAGE, type_name='.tabtest.Features')
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    return message_factory.GetMessageClass(pool.FindMessageTypeByName('tabtest.Example'))


def test_example_codec_against_protobuf():
    Example = _example_schema()
    rng = np.random.default_rng(11)
    blob = rng.normal(size=37).tobytes()
    ours = tfr.Example({
        'positions': tfr.bytes_feature(blob),
        'n_atoms_vap': tfr.int64_feature(12345678901),
        'neg': tfr.Feature('int64_list', [-1, 0, 7, -(1 << 62)]),
        'w': tfr.Feature('float_list', [0.5, -2.25, 1e-3]),
        'empty': tfr.bytes_feature(b''),
        'g2.indices': tfr.bytes_feature(bytes(range(200))),
    })
    wire = ours.SerializeToString()
    # protobuf reads what we write
    theirs = Example.FromString(wire)
    fm = theirs.features.feature
    assert set(fm.keys()) == set(ours.features.keys())
    assert fm['positions'].bytes_list.value[0] == blob
    assert list(fm['n_atoms_vap'].int64_list.value) == [12345678901]
    assert list(fm['neg'].int64_list.value) == [-1, 0, 7, -(1 << 62)]
    assert np.allclose(list(fm['w'].float_list.value), [0.5, -2.25, 1e-3], rtol=1e-7)
    assert fm['empty'].bytes_list.value[0] == b''
    # we read what protobuf writes (its own field order / packing)
    back = tfr.Example.FromString(theirs.SerializeToString())
    assert set(back.features) == set(ours.features)
    for k in ours.features:
        a, b = ours.features[k], back.features[k]
        assert a.kind == b.kind
        if a.kind == 'float_list':
            assert np.allclose(a.value, b.value, rtol=1e-7)
        else:
            assert a.value == b.value
    # and our own round trip is exact
    again = tfr.Example.FromString(wire)
    assert again.features['positions'].value == [blob]
    assert again.SerializeToString() == wire


def _labelled(atoms, seed, stress=True):
    rng = np.random.default_rng(seed)
    atoms.info['energy'] = float(rng.normal())
    atoms.info['forces'] = rng.normal(size=(len(atoms), 3))
    if stress:
        atoms.info['stress'] = rng.normal(size=6) * 0.01
    atoms.info['etemperature'] = 0.17
    atoms.info['eentropy'] = 0.4
    return atoms


def _structures():
    a = bulk_fcc('Ni', 3.6, (2, 2, 1))
    rng = np.random.default_rng(7)
    sym = ['Mo' if x < 0.4 else 'Ni' for x in rng.random(len(a))]
    a = Atoms(sym, a.positions + rng.normal(scale=0.08, size=a.positions.shape), a.cell, True)
    b = bulk_fcc('Ni', 3.52, (2, 1, 1))
    b.positions += rng.normal(scale=0.05, size=b.positions.shape)
    c = Atoms(['Ni', 'Mo', 'Ni', 'Mo', 'Mo'], rng.random((5, 3)) * 3 + 1.0, np.eye(3) * 9.0, True)
    return [_labelled(a, 1), _labelled(b, 2), _labelled(c, 3)]


def _nl(atoms, rc):
    return onl.neighbor_list(atoms.positions, atoms.cell, atoms.pbc, rc)[:3]


@pytest.mark.parametrize('angular', [False, True])
@pytest.mark.parametrize('precision', ['high', 'medium'])
def test_encode_decode_structures(angular, precision):
    images = _structures()
    rc = 4.0
    lists = [_nl(a, rc) for a in images]
    nij_max = max(len(l[0]) for l in lists) + 3
    clf0 = BatchUniversalTransformer({'Mo': 10, 'Ni': 16}, rcut=rc, angular=angular)
    with precision_scope(precision):
        nijk_max = None
        if angular:
            nijk_max = max(clf0.get_metadata(a, neighbor_list=l)['g4.v2g_map'].shape[0]
                           for a, l in zip(images, lists)) + 5
        clf = BatchUniversalTransformer({'Mo': 10, 'Ni': 16}, rcut=rc, angular=angular,
                                        nij_max=nij_max, nijk_max=nijk_max, use_stress=True)
        dt = np.float64 if precision == 'high' else np.float32
        for atoms, nl in zip(images, lists):
            wire = clf.encode(atoms, neighbor_list=nl).SerializeToString()
            dec = clf.decode_protobuf(wire)
            n = len(atoms)
            vap = clf.get_dataset_vap(atoms)
            # base.py:383-437
            assert dec['positions'].shape == (27, 3) and dec['positions'].dtype == dt
            assert np.array_equal(dec['positions'], vap.map_positions(atoms.positions).astype(dt))
            assert dec['n_atoms_vap'] == n
            assert dec['atom_masks'].sum() == n and dec['atom_masks'][0] == 0
            assert np.array_equal(dec['cell'], atoms.cell.astype(dt))
            assert dec['volume'] == dt(atoms.get_volume())
            assert dec['energy'] == dt(atoms.info['energy'])
            assert np.isclose(dec['free_energy'], atoms.info['energy'] - 0.17 * 0.4, rtol=1e-6)
            assert np.array_equal(dec['forces'], vap.map_forces(atoms.info['forces']).astype(dt))
            assert np.array_equal(dec['stress'], atoms.info['stress'].astype(dt))
            assert np.isclose(dec['total_pressure'],
                              -atoms.info['stress'][:3].mean() * 160.21766208, rtol=1e-6)
            # universal.py:46-112 in TRAIN mode: 6 columns, batch column 0, zero padding, mask
            nij = len(nl[0])
            v2g = dec['g2.v2g_map']
            assert v2g.shape == (nij_max, 6) and v2g.dtype == np.int32
            assert not v2g[:, 0].any() and not v2g[nij:].any()
            assert v2g[:nij, 5].all() and not v2g[:, 4].any()
            assert (dec['g2.ilist'][:nij] > 0).all() and not dec['g2.ilist'][nij:].any()
            assert np.array_equal(v2g[:, 2], dec['g2.ilist'])
            assert dec['g2.n1'].shape == (nij_max, 3) and dec['g2.n1'].dtype == dt
            # the pairs are those of the list (as a multiset of (i, j, S) in GSL numbering)
            l2g = vap.local_to_gsl_array
            want = sorted(zip(l2g[nl[0] + 1].tolist(), l2g[nl[1] + 1].tolist(),
                              map(tuple, nl[2].tolist())))
            got = sorted(zip(dec['g2.ilist'][:nij].tolist(), dec['g2.jlist'][:nij].tolist(),
                             map(tuple, dec['g2.n1'][:nij].astype(int).tolist())))
            assert want == got
            if angular:
                g4 = dec['g4.v2g_map']
                assert g4.shape == (nijk_max, 6)
                nijk = int(g4[:, 5].sum())
                assert not g4[nijk:].any() and not g4[:, 0].any()
                assert np.array_equal(dec['g4.n3'][:nijk], dec['g4.n2'][:nijk] - dec['g4.n1'][:nijk])
            # the structure itself comes back
            back = clf.decode_atoms(dec)
            assert sorted(back.get_chemical_symbols()) == sorted(atoms.get_chemical_symbols())
            order = vap.local_to_gsl_array[1:] - 1      # local -> position among the real rows
            rank = np.argsort(np.argsort(order))
            tol = 0 if precision == 'high' else 1e-6
            assert np.allclose(back.positions[rank], atoms.positions, atol=tol * 10, rtol=tol)
            assert np.allclose(back.info['forces'][rank], atoms.info['forces'], atol=tol * 10,
                               rtol=tol)
        # a record of another size is refused (tf: set_shape)
        small = BatchUniversalTransformer({'Mo': 10, 'Ni': 16}, rcut=rc, angular=angular,
                                          nij_max=nij_max + 1, nijk_max=nijk_max,
                                          use_stress=True)
        with pytest.raises(ValueError):
            small.decode_protobuf(wire)
        tight = BatchUniversalTransformer({'Mo': 10, 'Ni': 16}, rcut=rc, nij_max=5)
        with pytest.raises(ValueError):
            tight.encode(images[0], neighbor_list=lists[0])


def test_dataset_record_files(tmp_path):
    images = _structures() + [_labelled(bulk_fcc('Ni', 3.5, (1, 1, 2)), 9)]
    rc = 3.8
    lists = [_nl(a, rc) for a in images]
    ds = Dataset.from_images(images)
    clf = BatchUniversalTransformer(ds.max_occurs, rcut=rc, use_stress=True,
                                    nij_max=max(len(l[0]) for l in lists))
    with precision_scope('high'):
        files = ds.to_records(str(tmp_path), clf, name='moni', test_size=[2],
                              neighbor_lists=lists)
        assert files['test'].endswith('moni-test-k2-rc3.80-fp64-1.universal.tfrecords')
        assert files['train'].endswith('moni-train-k2-rc3.80-fp64-3.universal.tfrecords')
        test = Dataset.from_records(files['test'], clf)
        train = Dataset.from_records(files['train'], clf)
    assert len(test) == 1 and len(train) == 3
    assert test[0].info['energy'] == images[1].info['energy']
    assert [a.info['energy'] for a in train] == [images[k].info['energy'] for k in (0, 2, 3)]
    assert train.has_stress and train.max_occurs['Ni'] <= ds.max_occurs['Ni']
