"""Attribute-bag DTO (API of the reference's common/dto/Dto.py:1-44)."""


class Dto():
    """Data transfer object: keyword arguments become attributes; iterable as (name, value) pairs."""

    def __init__(self, **kwargs):
        self.__dict__ = kwargs

    def __iter__(self):
        for item in self.__dict__.items():
            yield item

    def __str__(self, indent=None):
        """Fill-level listing: ``[x] name`` for set attributes, ``[ ] name`` for None, nested DTOs indented."""
        lines = []
        if indent is None:
            lines.append('Fill level of ' + object.__str__(self) + ':')
            indent = ''
        for name in sorted(self.__dict__):
            value = self.__dict__[name]
            lines.append('%s[%s] %s' % (indent, ' ' if value is None else 'x', name))
            if isinstance(value, Dto):
                nested = value.__str__(indent=indent + '    ')
                if nested:
                    lines.append(nested.rstrip('\n'))
        return '\n'.join(lines) + '\n'

    def _is_empty(self):
        """True when no direct non-DTO attribute is set.

        Mirrors the reference (Dto.py:36-44), which evaluates nested DTOs but drops their result; the models' guard
        asserts (Cae3D.py:104,112,229,235) depend on exactly that behaviour.
        """
        for value in self.__dict__.values():
            if value is not None and not isinstance(value, Dto):
                return False
        return True
