"""R-fmt / R-match -- CPU restatement of MatcherParser's header extraction and `<*>` template
matching, done with Python's regular-expression engine.  TEST INFRASTRUCTURE.

PARITY UNPINNED in general: the class is detectmatelibrary.parsers.template_matcher.MatcherParser
(detectmatelibrary 0.1.0 @ ecdda558, /root/reference/uv.lock:240-251), absent from
/root/reference.  What the reference does show: its configuration
(tests/library_integration/test_pipe_filereader_matcher_nvd.py:74-88: log_format
"type=<type> msg=audit(<Time>...): <Content>", templates file audit_templates.txt;
docs/getting_started.md:395-415: the nginx access-log log_format) and one end-to-end known
answer: three access-log requests /hello, /world, /foobar with data_use_training 2 alert
on the third with {"Global - URL": "Unknown value: '/foobar'"} (docs/getting_started.md:423-435,
498-510) -- pinned in tests/test_format_matcher.py.

R-fmt   log_format = literal text with <Name> captures (Name = [A-Za-z0-9_]+).  A record
        matches iff  ^L0(.*?)L1(.*?)...$  matches it, literals escaped, captures non-greedy,
        '.' matching any byte.  Captures are the record's header variables
        (logFormatVariables[Name]).  A record that does not match has no fields and never
        alerts.  Two captures with nothing between them are a configuration error.
R-match templates = lines of the templates file; each is literal text with `<*>` wildcards and
        matches the capture named Content iff ^S0(.*?)S1(.*?)...$ matches it.  The first
        matching template (file order) gives EventID = its 0-based index and variables =
        its wildcard captures; no match: EventID -1, no variables.
R-norm  params.remove_spaces / remove_punctuation / lowercase (all true in the reference's audit
        config, test_pipe_filereader_matcher_nvd.py:82-84; all false in docs/getting_started.md:
        411-413 and container/config/parser_config.yaml:8-10).  The reference shows the switches
        but not their effect, so this is a restatement of the library's documented intent, PARITY
        UNPINNED: before template matching, the Content text AND every literal segment of every
        template (the `<*>` wildcards survive) are normalised byte-wise --
          remove_spaces       drop 0x09-0x0D and 0x20            (ASCII white space)
          remove_punctuation  drop the 32 bytes of string.punctuation
          lowercase           'A'-'Z' -> 'a'-'z'                  (ASCII only; other bytes untouched)
        -- and R-match then runs on the normalised text: `variables` are slices of the normalised
        Content.  The header extraction (R-fmt) and logFormatVariables stay verbatim.  A template
        whose normalisation leaves nothing between two wildcards is a configuration error.
"""
from __future__ import annotations

import re
import string
from typing import Dict, List, Optional, Tuple

_CAPTURE = re.compile(rb"<([A-Za-z0-9_]+)>")

_SPACES = bytes(range(0x09, 0x0E)) + b" "
_PUNCT = string.punctuation.encode("ascii")
_UPPER = bytes(range(0x41, 0x5B))
_LOWER = bytes(range(0x61, 0x7B))


def normaliser(remove_spaces: bool = False, remove_punctuation: bool = False, lowercase: bool = False):
    """R-norm as a bytes -> bytes function."""
    drop = (_SPACES if remove_spaces else b"") + (_PUNCT if remove_punctuation else b"")
    table = bytes.maketrans(_UPPER, _LOWER) if lowercase else None
    if not drop and table is None:
        return lambda text: text
    return lambda text: text.translate(table, drop)


def _chain_regex(literals: List[bytes], ends_with_capture: bool) -> "re.Pattern[bytes]":
    src = b"^"
    for i, lit in enumerate(literals):
        if i:
            src += b"(.*?)"
        src += re.escape(lit)
    if ends_with_capture:
        src += b"(.*?)"
    return re.compile(src + b"$", re.DOTALL)


def compile_log_format(fmt: bytes) -> Tuple["re.Pattern[bytes]", List[str]]:
    parts = _CAPTURE.split(fmt)                       # lit0, name0, lit1, name1, ..., litN
    lits, names = parts[0::2], [n.decode() for n in parts[1::2]]
    if any(l == b"" for l in lits[1:-1]):
        raise ValueError("two captures with nothing between them")
    if len(set(names)) != len(names):
        raise ValueError("a capture name appears twice")
    ends = bool(names) and lits[-1] == b""
    return _chain_regex(lits[:-1] if ends else lits, ends), names


def compile_template(t: bytes, norm=lambda x: x) -> "re.Pattern[bytes]":
    segs = [norm(s) for s in t.split(b"<*>")]
    if any(s == b"" for s in segs[1:-1]):
        raise ValueError("two wildcards with nothing between them")
    ends = len(segs) > 1 and segs[-1] == b""
    return _chain_regex(segs[:-1] if ends else segs, ends)


class FormatParser:
    def __init__(self, log_format, templates=(), content_name: str = "Content", remove_spaces: bool = False,
                 remove_punctuation: bool = False, lowercase: bool = False) -> None:
        b = lambda x: x if isinstance(x, bytes) else str(x).encode("utf-8")
        self.norm = normaliser(remove_spaces, remove_punctuation, lowercase)
        self.regex, self.names = compile_log_format(b(log_format))
        self.templates = [compile_template(b(t), self.norm) for t in templates]
        self.content_name = content_name

    def parse_line(self, line: bytes) -> Optional[dict]:
        """One record (without its '\\n') -> ParserSchema-like dict, or None if the
        log_format does not match."""
        m = self.regex.match(line)
        if m is None:
            return None
        lfv: Dict[str, bytes] = {n: m.group(i + 1) for i, n in enumerate(self.names)}
        eid, variables = -1, []
        content = lfv.get(self.content_name)
        if content is not None:
            content = self.norm(content)
            for t, rx in enumerate(self.templates):
                tm = rx.match(content)
                if tm is not None:
                    eid, variables = t, list(tm.groups())
                    break
        return {"EventID": eid, "variables": variables, "logFormatVariables": lfv}
