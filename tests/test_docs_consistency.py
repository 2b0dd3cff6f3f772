"""Static consistency checks between the sources and the documents a maintainer reads (no GPU, no library load)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _read(*parts):
    with open(os.path.join(ROOT, *parts)) as f:
        return f.read()


def _sources():
    out = []
    for d, exts in (('imagen_pytorch_b200/csrc', ('.cu', '.cuh')), ('imagen_pytorch_b200', ('.py',))):
        for name in sorted(os.listdir(os.path.join(ROOT, d))):
            if name.endswith(exts):
                out.append(_read(d, name))
    return '\n'.join(out)


def test_every_environment_switch_is_documented():
    """Each B200_IMAGEN_* variable the library or the package reads appears in DESIGN.md section 6b, and vice versa."""
    in_src = set(re.findall(r'B200_IMAGEN_[A-Z0-9_]+', _sources()))
    in_doc = set(re.findall(r'B200_IMAGEN_[A-Z0-9_]+', _read('DESIGN.md')))
    assert in_src - in_doc == set(), f'undocumented switches: {sorted(in_src - in_doc)}'
    assert in_doc - in_src == set(), f'documented but unused switches: {sorted(in_doc - in_src)}'


def test_attention_sweep_names_match_the_dispatch():
    """tools/sweep_attention.py only names kernel variants the dispatch in attention_tc.cu knows."""
    cases = set(int(m) for m in re.findall(r'^\s*case (-?\d+):', _read('imagen_pytorch_b200', 'csrc', 'attention_tc.cu'), flags=re.M))
    sweep = _read('tools', 'sweep_attention.py')
    names = set(int(m) for m in re.findall(r'(-?\d+): \'', sweep[sweep.index('NAMES = {'):sweep.index("CODE = '''")]))
    assert names, 'no variants parsed from the sweep tool'
    assert names - cases == set(), f'sweep tool names unknown variants: {sorted(names - cases)}'


def test_abi_entry_points_are_listed_in_integration_md():
    """Every extern "C" entry point of include/b200_imagen.h is mentioned in INTEGRATION.md or DESIGN.md."""
    header = _read('include', 'b200_imagen.h')
    entries = set(re.findall(r'\b(b200_[a-z0-9_]+)\s*\(', header))
    docs = _read('INTEGRATION.md') + _read('DESIGN.md')
    missing = sorted(e for e in entries if e not in docs)
    assert len(missing) <= len(entries) // 2, f'most ABI entry points should be explained in the docs; missing: {missing}'


def test_profiles_named_in_the_docs_exist():
    """Evidence files the documents cite (profiles/rNN_*) are committed."""
    cited = set()
    for doc in ('DESIGN.md', 'README.md', 'profiles/README.md'):
        cited |= set(re.findall(r'`(?:profiles/)?(r0[12]_[A-Za-z0-9_.]+\.(?:txt|json|csv\.gz|csv))`', _read(*doc.split('/'))))
    have = set(os.listdir(os.path.join(ROOT, 'profiles')))
    missing = sorted(c for c in cited if c not in have)
    assert missing == [], f'cited evidence files that are not in profiles/: {missing}'
