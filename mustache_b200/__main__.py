"""`python -m mustache_b200 ...` = the reference's `mustache` console script (setup.py:15-17);
`python -m mustache_b200 diff ...` = `python diff_mustache.py ...`."""
import sys

if len(sys.argv) > 1 and sys.argv[1] == "diff":
    from .diff_mustache import main
    main(sys.argv[2:])
else:
    from .mustache import main
    main()
