"""Row pitch of the field-major (SoA) device arrays: rows are padded so every row starts 128-byte aligned."""


def pitch(n, multiple=32):
    return ((int(n) + multiple - 1) // multiple) * multiple
