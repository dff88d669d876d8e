"""Error type that reproduces the reference's `print(msg); sys.exit()` behaviour.

The reference reports user errors on this path by printing a message and calling
sys.exit() (MAS_library.pyx:64-66,82,109; Pk_library.pyx:95-99,560).  `ReferenceExit`
is a SystemExit (so an unguarded script ends exactly like it does with the reference, exit
status 0 after the message has been printed) AND a ValueError (so library users and tests
can catch it like a normal argument error)."""


class ReferenceExit(SystemExit, ValueError):
    def __init__(self, message=""):
        SystemExit.__init__(self)          # code=None -> exit status 0, like sys.exit()
        self.message = message

    def __str__(self):
        return self.message


def reference_exit(*lines):
    for ln in lines:
        print(ln)
    raise ReferenceExit("\n".join(lines))
