"""Minimal OOXML workbook writer / reader (openpyxl is not available in this image).

`yacht run` writes results/result.xlsx with one sheet per min_coverage (reference
src/yacht/run_YACHT.py:231-254, pandas.ExcelWriter(engine="openpyxl")).  This module writes the same
logical content -- sheet names, header row, cell values and types (inline strings, numbers,
booleans) -- with nothing but zipfile; `read_xlsx` reads such workbooks (and openpyxl's) back.
"""
from __future__ import annotations

import re
import zipfile
from typing import Dict, List, Sequence
from xml.sax.saxutils import escape

import numpy as np
import pandas as pd


def _col(k: int) -> str:
    s = ""
    k += 1
    while k:
        k, r = divmod(k - 1, 26)
        s = chr(65 + r) + s
    return s


def _cell(ref: str, v, header: bool = False) -> str:
    if v is None or (isinstance(v, float) and np.isnan(v)):
        return ""
    if isinstance(v, (bool, np.bool_)):
        return f'<c r="{ref}" t="b"><v>{int(v)}</v></c>'
    if isinstance(v, (int, np.integer)):
        return f'<c r="{ref}" t="n"><v>{int(v)}</v></c>'
    if isinstance(v, (float, np.floating)):
        return f'<c r="{ref}" t="n"><v>{repr(float(v))}</v></c>'
    style = ' s="1"' if header else ""
    return f'<c r="{ref}"{style} t="inlineStr"><is><t>{escape(str(v))}</t></is></c>'


def _sheet_xml(df: pd.DataFrame) -> str:
    rows = []
    cols = list(df.columns)
    rows.append('<row r="1">' + "".join(_cell(f"{_col(c)}1", name, True) for c, name in enumerate(cols)) + "</row>")
    for r, rec in enumerate(df.itertuples(index=False, name=None), start=2):
        rows.append(f'<row r="{r}">' + "".join(_cell(f"{_col(c)}{r}", v) for c, v in enumerate(rec)) + "</row>")
    dim = f"A1:{_col(max(len(cols) - 1, 0))}{len(df) + 1}"
    return ('<?xml version="1.0" encoding="UTF-8" standalone="yes"?>'
            '<worksheet xmlns="http://schemas.openxmlformats.org/spreadsheetml/2006/main">'
            f'<dimension ref="{dim}"/><sheetData>' + "".join(rows) + "</sheetData></worksheet>")


_CT = "application/vnd.openxmlformats-officedocument.spreadsheetml"
_REL = "http://schemas.openxmlformats.org/officeDocument/2006/relationships"


def write_xlsx(path: str, sheets: Sequence) -> None:
    """sheets: [(sheet name, DataFrame), ...] in workbook order."""
    names = [n for n, _ in sheets]
    k = len(names)
    with zipfile.ZipFile(path, "w", zipfile.ZIP_DEFLATED) as z:
        z.writestr("[Content_Types].xml",
                   '<?xml version="1.0" encoding="UTF-8" standalone="yes"?>'
                   '<Types xmlns="http://schemas.openxmlformats.org/package/2006/content-types">'
                   '<Default Extension="rels" ContentType="application/vnd.openxmlformats-package.relationships+xml"/>'
                   '<Default Extension="xml" ContentType="application/xml"/>'
                   f'<Override PartName="/xl/workbook.xml" ContentType="{_CT}.sheet.main+xml"/>'
                   f'<Override PartName="/xl/styles.xml" ContentType="{_CT}.styles+xml"/>'
                   + "".join(f'<Override PartName="/xl/worksheets/sheet{i + 1}.xml" ContentType="{_CT}.worksheet+xml"/>' for i in range(k))
                   + "</Types>")
        z.writestr("_rels/.rels",
                   '<?xml version="1.0" encoding="UTF-8" standalone="yes"?>'
                   '<Relationships xmlns="http://schemas.openxmlformats.org/package/2006/relationships">'
                   f'<Relationship Id="rId1" Type="{_REL}/officeDocument" Target="xl/workbook.xml"/></Relationships>')
        z.writestr("xl/workbook.xml",
                   '<?xml version="1.0" encoding="UTF-8" standalone="yes"?>'
                   '<workbook xmlns="http://schemas.openxmlformats.org/spreadsheetml/2006/main" '
                   f'xmlns:r="{_REL}"><sheets>'
                   + "".join(f'<sheet name="{escape(n, {chr(34): "&quot;"})}" sheetId="{i + 1}" r:id="rId{i + 1}"/>' for i, n in enumerate(names))
                   + "</sheets></workbook>")
        z.writestr("xl/_rels/workbook.xml.rels",
                   '<?xml version="1.0" encoding="UTF-8" standalone="yes"?>'
                   '<Relationships xmlns="http://schemas.openxmlformats.org/package/2006/relationships">'
                   + "".join(f'<Relationship Id="rId{i + 1}" Type="{_REL}/worksheet" Target="worksheets/sheet{i + 1}.xml"/>' for i in range(k))
                   + f'<Relationship Id="rId{k + 1}" Type="{_REL}/styles" Target="styles.xml"/></Relationships>')
        z.writestr("xl/styles.xml",
                   '<?xml version="1.0" encoding="UTF-8" standalone="yes"?>'
                   '<styleSheet xmlns="http://schemas.openxmlformats.org/spreadsheetml/2006/main">'
                   '<fonts count="2"><font><sz val="11"/><name val="Calibri"/></font><font><b/><sz val="11"/><name val="Calibri"/></font></fonts>'
                   '<fills count="1"><fill><patternFill patternType="none"/></fill></fills>'
                   '<borders count="1"><border/></borders><cellStyleXfs count="1"><xf/></cellStyleXfs>'
                   '<cellXfs count="2"><xf fontId="0"/><xf fontId="1" applyFont="1"/></cellXfs></styleSheet>')
        for i, (_, df) in enumerate(sheets):
            z.writestr(f"xl/worksheets/sheet{i + 1}.xml", _sheet_xml(df))


_ROW_RE = re.compile(r"<row [^>]*>(.*?)</row>", re.S)
_CELL_RE = re.compile(r'<c r="([A-Z]+)\d+"(?: s="\d+")?(?: t="(\w+)")?\s*(?:/>|>(.*?)</c>)', re.S)
_VAL_RE = re.compile(r"<v>(.*?)</v>|<t[^>]*>(.*?)</t>", re.S)


def _unescape(s: str) -> str:
    return s.replace("&lt;", "<").replace("&gt;", ">").replace("&quot;", '"').replace("&apos;", "'").replace("&amp;", "&")


def read_xlsx(path: str) -> Dict[str, pd.DataFrame]:
    """{sheet name: DataFrame} (first row = header)."""
    out: Dict[str, pd.DataFrame] = {}
    with zipfile.ZipFile(path) as z:
        wb = z.read("xl/workbook.xml").decode()
        names = re.findall(r'<sheet [^>]*name="([^"]*)"', wb)
        shared: List[str] = []
        if "xl/sharedStrings.xml" in z.namelist():
            shared = [_unescape(re.sub(r"<[^>]+>", "", m))
                      for m in re.findall(r"<si>(.*?)</si>", z.read("xl/sharedStrings.xml").decode(), re.S)]
        for i, name in enumerate(names):
            xml = z.read(f"xl/worksheets/sheet{i + 1}.xml").decode()
            table = []
            for rm in _ROW_RE.finditer(xml):
                cells = {}
                for col, typ, body in _CELL_RE.findall(rm.group(1)):
                    if not body:
                        continue
                    vm = _VAL_RE.search(body)
                    if not vm:
                        continue
                    txt = vm.group(1) if vm.group(1) is not None else vm.group(2)
                    if typ == "s":
                        v = shared[int(txt)]
                    elif typ in ("inlineStr", "str"):
                        v = _unescape(txt)
                    elif typ == "b":
                        v = bool(int(txt))
                    else:
                        f = float(txt)
                        v = int(f) if (f.is_integer() and "." not in txt and "e" not in txt.lower()) else f
                    cells[col] = v
                table.append(cells)
            if not table:
                out[_unescape(name)] = pd.DataFrame()
                continue
            hdr_cols = sorted(table[0].keys(), key=lambda c: (len(c), c))
            header = [table[0][c] for c in hdr_cols]
            out[_unescape(name)] = pd.DataFrame([[r.get(c) for c in hdr_cols] for r in table[1:]], columns=header)
    return out
