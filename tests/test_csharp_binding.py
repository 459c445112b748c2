"""Static cross-check of the reference-side binding (csharp/JpegLibrary.Cuda/Native.cs) against include/jpegb200.h.

No .NET toolchain exists in this image, so the P/Invoke stub is never compiled here; what CAN drift silently -- entry
point names, parameter counts, status / format constants and the layouts of the blittable structs -- is checked
from the sources: the C side by compiling a probe with gcc (sizeof / offsetof of the header's structs), the C# side by
laying the `[StructLayout(LayoutKind.Sequential)]` structs out with the CLR's rules for blittable fields (natural
alignment, `fixed` buffers as inline arrays), the Python side (jpeglibrary_b200/_native.py) through ctypes.
"""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NATIVE_CS = os.path.join(ROOT, "csharp", "JpegLibrary.Cuda", "Native.cs")
HEADER = os.path.join(ROOT, "include", "jpegb200.h")

CS_TYPES = {"byte": 1, "sbyte": 1, "short": 2, "ushort": 2, "int": 4, "uint": 4, "long": 8, "ulong": 8, "float": 4,
            "double": 8, "IntPtr": 8, "UIntPtr": 8}

# C# struct -> (C struct, ctypes mirror)
STRUCTS = {"HuffSpec": "jb_huff_spec", "ScanDesc": "jb_scan_desc", "ImageDesc": "jb_image_desc",
           "OutputDesc": "jb_output_desc", "EncodeDesc": "jb_encode_desc"}


def _strip_comments(text):
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    return re.sub(r"//[^\n]*", "", text)


def _cs_source():
    return _strip_comments(open(NATIVE_CS).read())


def _header_source():
    return _strip_comments(open(HEADER).read())


def _cs_imports():
    out = {}
    for m in re.finditer(r"\[DllImport\(Lib\)\]\s*public\s+static\s+extern\s+([\w\*]+)\s+(\w+)\s*\(([^;]*?)\)\s*;", _cs_source(), flags=re.S):
        args = m.group(3).strip()
        out[m.group(2)] = (m.group(1), [] if not args else [a.strip() for a in args.split(",")])
    return out


def _c_prototypes():
    out = {}
    for m in re.finditer(r"JB_API\s+([\w\s\*]+?)\b(jb_\w+)\s*\(([^;]*?)\)\s*;", _header_source(), flags=re.S):
        args = " ".join(m.group(3).split())
        out[m.group(2)] = (" ".join(m.group(1).split()), [] if args in ("", "void") else [a.strip() for a in args.split(",")])
    return out


def _cs_struct_layouts():
    """name -> (size, [(field, offset, size)]) for every sequential struct of Native.cs."""
    src = _cs_source()
    layouts = {}
    for m in re.finditer(r"\[StructLayout\(LayoutKind\.Sequential\)\]\s*public\s+struct\s+(\w+)\s*\{(.*?)\}", src, flags=re.S):
        name, body = m.group(1), m.group(2)
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            dm = re.match(r"public\s+(fixed\s+)?([\w\.]+\*?)\s+(.*)$", decl, flags=re.S)
            assert dm, f"{name}: cannot parse field declaration {decl!r}"
            fixed, typ, names = bool(dm.group(1)), dm.group(2), dm.group(3)
            for item in names.split(","):
                item = item.strip()
                am = re.match(r"(\w+)\s*\[([^\]]+)\]$", item)
                if fixed:
                    assert am, f"{name}: fixed buffer without a length: {item!r}"
                    count = eval(am.group(2), {"__builtins__": {}})  # "4 * 64"
                    fields.append((am.group(1), CS_TYPES[typ], CS_TYPES[typ] * count))
                else:
                    assert not am, f"{name}: array field that is not a fixed buffer: {item!r}"
                    size = 8 if typ.endswith("*") else CS_TYPES[typ]
                    fields.append((item, size, size))
        off, align, placed = 0, 1, []
        for fname, falign, fsize in fields:
            off = (off + falign - 1) // falign * falign
            placed.append((fname, off, fsize))
            off += fsize
            align = max(align, falign)
        layouts[name] = ((off + align - 1) // align * align, placed)
    return layouts


def _c_struct_layouts(tmp_path):
    """The same from the C compiler: a probe program printing sizeof and the offset of every member."""
    hdr = _header_source()
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', 'int main(void) {']
    for cname in STRUCTS.values():
        body = re.search(r"typedef\s+struct\s+" + cname + r"\s*\{(.*?)\}\s*" + cname + r"\s*;", hdr, flags=re.S).group(1)
        members = []
        for decl in body.split(";"):
            decl = " ".join(decl.split())
            if not decl:
                continue
            for item in decl.split(","):
                members.append(re.match(r".*?(\w+)\s*(?:\[[^\]]*\]\s*)*$", item.strip()).group(1))
        lines.append(f'  printf("{cname} %zu", sizeof({cname}));')
        for mname in members:
            lines.append(f'  printf(" %zu:%zu", offsetof({cname}, {mname}), sizeof((({cname} *)0)->{mname}));')
        lines.append('  printf("\\n");')
    lines += ['  return 0;', '}']
    src = tmp_path / "probe.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "probe"
    subprocess.run(["gcc", "-std=c11", "-o", str(exe), str(src)], check=True)
    out = {}
    for line in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines():
        parts = line.split()
        out[parts[0]] = (int(parts[1]), [tuple(int(v) for v in p.split(":")) for p in parts[2:]])
    return out


def test_every_dllimport_is_a_declared_entry_point_with_the_same_parameter_count():
    imports, protos = _cs_imports(), _c_prototypes()
    assert len(imports) >= 30
    for name, (_ret, args) in imports.items():
        assert name in protos, f"Native.cs imports {name}, which include/jpegb200.h does not declare"
        assert len(args) == len(protos[name][1]), f"{name}: C# passes {args}, the header takes {protos[name][1]}"


def test_dllimport_parameter_kinds_match_the_header():
    """Pointers stay pointers (IntPtr, T*, out IntPtr), integers keep their width."""
    width = {"int": 4, "uint": 4, "ulong": 8, "UIntPtr": 8, "int32_t": 4, "uint32_t": 4, "uint64_t": 8, "size_t": 8}
    for name, (_ret, args) in _cs_imports().items():
        for cs_arg, c_arg in zip(args, _c_prototypes()[name][1]):
            cs_type = cs_arg.rsplit(" ", 1)[0].strip()
            c_is_pointer = "*" in c_arg or "[" in c_arg  # an array parameter is a pointer
            cs_is_pointer = cs_type.endswith("*") or cs_type in ("IntPtr", "out IntPtr")
            assert cs_is_pointer == c_is_pointer, f"{name}: {cs_arg!r} against {c_arg!r}"
            if not c_is_pointer:
                c_type = c_arg.replace("const ", "").rsplit(" ", 1)[0].strip()
                assert width[cs_type] == width[c_type], f"{name}: {cs_arg!r} against {c_arg!r}"


def test_dllimports_are_exported_by_the_built_library():
    lib = os.path.join(ROOT, "jpeglibrary_b200", "lib", "libjpegb200.so")
    if not os.path.exists(lib):
        pytest.skip("library not built")
    exported = set(re.findall(r"\b(jb_\w+)\b", subprocess.run(["nm", "-D", "--defined-only", lib], check=True, capture_output=True, text=True).stdout))
    missing = [n for n in _cs_imports() if n not in exported]
    assert not missing, missing


def test_status_and_format_constants_equal_the_header():
    hdr = open(HEADER).read()
    defines = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define\s+(JB_\w+)\s+\(?(-?\d+)\)?", hdr)}
    consts = {}
    for m in re.finditer(r"public\s+const\s+int\s+([^;]+);", _cs_source(), flags=re.S):
        for item in m.group(1).split(","):
            k, v = item.split("=")
            consts[k.strip()] = int(v)
    assert len(consts) >= 13
    for k, v in consts.items():
        assert defines.get(k) == v, f"{k}: Native.cs says {v}, the header {defines.get(k)}"


def test_struct_layouts_agree_between_csharp_c_and_ctypes(tmp_path):
    from jpeglibrary_b200 import _native as N
    mirrors = {"HuffSpec": N.HuffSpec, "ScanDesc": N.ScanDesc, "ImageDesc": N.ImageDesc, "OutputDesc": N.OutputDesc,
               "EncodeDesc": N.EncodeDesc}
    cs, c = _cs_struct_layouts(), _c_struct_layouts(tmp_path)
    for cs_name, c_name in STRUCTS.items():
        cs_size, cs_fields = cs[cs_name]
        c_size, c_fields = c[c_name]
        assert cs_size == c_size, f"{cs_name}: {cs_size} bytes in C#, {c_size} in C"
        # C# may declare several C members as one run or the other way round (reserved bytes): compare the byte map
        assert [(o, s) for _n, o, s in cs_fields] == c_fields, f"{cs_name}: member offsets differ: {cs_fields} against {c_fields}"
        py = mirrors[cs_name]
        assert C.sizeof(py) == c_size
        assert [(getattr(py, f[0]).offset, getattr(py, f[0]).size) for f in py._fields_] == c_fields


def test_every_native_member_the_binding_uses_is_declared():
    import glob
    nat = _cs_source()
    declared = (set(re.findall(r"extern\s+[\w\*]+\s+(\w+)\s*\(", nat)) | set(re.findall(r"\b(JB_\w+)\s*=", nat))
                | set(re.findall(r"struct\s+(\w+)", nat)) | {"Check"})
    files = glob.glob(os.path.join(ROOT, "csharp", "JpegLibrary.Cuda", "*.cs"))
    assert len(files) >= 6
    for path in files:
        used = set(re.findall(r"Native\.(\w+)", _strip_comments(open(path).read())))
        assert not used - declared, f"{os.path.basename(path)} uses undeclared Native members {sorted(used - declared)}"
