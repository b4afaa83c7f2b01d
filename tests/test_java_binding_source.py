"""The Java binding (bindings/java/.../JaicovB200.java) cannot be compiled here (no JDK).  What CAN be checked:
every downcall descriptor it declares matches the prototype of include/jaicov_b200.h -- symbol exists, same number of arguments,
same FFM layout per argument (int32_t -> JAVA_INT, int64_t -> JAVA_LONG, double -> JAVA_DOUBLE, pointers / callbacks -> ADDRESS)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
JAVA = os.path.join(ROOT, 'bindings', 'java', 'org', 'applied_geodesy', 'adjustment', 'bundle', 'gpu', 'JaicovB200.java')


def c_layout(ctype):
    t = ctype.strip()
    if '*' in t or '[' in t or t.startswith('jaicov_progress_cb'):
        return 'ADDRESS'
    t = re.sub(r'\b(const|volatile)\b', '', t).split()
    base = t[0]
    return {'int32_t': 'JAVA_INT', 'int64_t': 'JAVA_LONG', 'double': 'JAVA_DOUBLE', 'void': 'VOID'}[base]


def header_prototypes():
    src = open(os.path.join(ROOT, 'include', 'jaicov_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    protos = {}
    for m in re.finditer(r'([A-Za-z_][A-Za-z0-9_ ]*?[\s\*]+)(jaicov_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;', src):
        ret, name, args = m.group(1), m.group(2), m.group(3)
        if 'typedef' in ret:
            continue
        args = [a for a in (x.strip() for x in args.replace('\n', ' ').split(',')) if a and a != 'void']
        protos[name] = (c_layout(ret), [c_layout(a) for a in args])
    return protos


def descriptors(text):
    out = {}
    for m in re.finditer(r'h\("(jaicov_[a-z0-9_]+)",\s*FunctionDescriptor\.(ofVoid|of)\(([^;]*?)\)\)\s*;', text):
        name, kind, args = m.group(1), m.group(2), [a.strip() for a in m.group(3).split(',') if a.strip()]
        out[name] = ('VOID', args) if kind == 'ofVoid' else (args[0], args[1:])
    return out


def test_header_parser_sees_every_entry_point():
    protos = header_prototypes()
    assert len(protos) == 41
    assert protos['jaicov_destroy'] == ('VOID', ['ADDRESS'])
    assert protos['jaicov_last_error'] == ('ADDRESS', ['ADDRESS'])
    assert protos['jaicov_get_qxx_block'] == ('JAVA_INT', ['ADDRESS', 'JAVA_INT', 'JAVA_INT', 'JAVA_INT', 'JAVA_INT', 'ADDRESS', 'JAVA_LONG'])
    assert protos['jaicov_release_cached_memory'] == ('JAVA_LONG', [])


@pytest.mark.parametrize('where', ['bindings/java', 'INTEGRATION.md'])
def test_java_downcall_descriptors_match_the_header(where):
    text = open(JAVA if where == 'bindings/java' else os.path.join(ROOT, 'INTEGRATION.md')).read()
    desc = descriptors(text)
    assert len(desc) >= 13
    protos = header_prototypes()
    for name, d in desc.items():
        assert name in protos, name
        assert d == protos[name], (name, d, protos[name])


def test_binding_covers_the_call_sequence_of_estimate_model():
    desc = descriptors(open(JAVA).read())
    for name in ('jaicov_create', 'jaicov_set_cameras', 'jaicov_set_images', 'jaicov_set_image_points', 'jaicov_set_object_points',
                 'jaicov_set_scale_bars', 'jaicov_add_observed_group', 'jaicov_set_datum', 'jaicov_set_reduced_rows', 'jaicov_estimate',
                 'jaicov_get_values', 'jaicov_get_stats', 'jaicov_get_qxx_packed', 'jaicov_get_qxx_block', 'jaicov_get_qxx_submatrix',
                 'jaicov_last_error', 'jaicov_destroy'):
        assert name in desc, name


def test_ctypes_binding_matches_the_header(built):
    """Same check for the ctypes binding the parity tests and bench.py go through (bundle-adjustment_b200/_lib.py): a c_int where the
    header says int64_t would pass small tests and corrupt large ones."""
    import ctypes

    import bundle_adjustment_b200 as ba
    L = ba._lib.load()
    protos = header_prototypes()

    def layout(t):
        if t is None:
            return 'VOID'
        if t in (ctypes.c_int32, ctypes.c_int):
            return 'JAVA_INT'
        if t in (ctypes.c_int64, ctypes.c_longlong, ctypes.c_long):
            return 'JAVA_LONG' if ctypes.sizeof(t) == 8 else 'JAVA_INT'
        if t is ctypes.c_double:
            return 'JAVA_DOUBLE'
        return 'ADDRESS'                                     # c_void_p, c_char_p, POINTER(...), CFUNCTYPE

    for name, (ret, args) in protos.items():
        f = getattr(L, name)
        assert f.argtypes is not None, '%s: argtypes not declared' % name
        assert [layout(t) for t in f.argtypes] == args, (name, f.argtypes, args)
        assert layout(f.restype) == ret, (name, f.restype, ret)


def test_integration_excerpt_uses_the_names_of_the_shipped_java_sources():
    """INTEGRATION.md section 3 patches estimateModel() with JaicovB200.<HANDLE> and FlatProblem fields: they exist in bindings/java."""
    doc = open(os.path.join(ROOT, 'INTEGRATION.md')).read()
    java = open(JAVA).read()
    flat = open(os.path.join(os.path.dirname(JAVA), 'FlatProblem.java')).read()
    handles = set(re.findall(r'\bJaicovB200\.([A-Z_]+)\b', doc))
    assert len(handles) >= 12
    assert not [h for h in handles if not re.search(r'MethodHandle\s+%s\b' % h, java)]
    fields = set(re.findall(r'\bf\.(\w+)', doc))
    assert len(fields) >= 25
    assert not [f for f in fields if not re.search(r'\b%s\b' % f, flat)]
