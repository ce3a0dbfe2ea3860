"""``hypha-certutil {root,org,node}`` (crates/certutil/src/cli.rs:59-279) plus ``crl`` (revocation lists)."""
import argparse
import sys

from .. import certutil as cu


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(prog="hypha-certutil", description="development PKI for hypha nodes")
    sub = p.add_subparsers(dest="command", required=True)
    r = sub.add_parser("root", help="self-signed Ed25519 root CA")
    r.add_argument("-o", "--organization", required=True)
    r.add_argument("--country", default="US")
    r.add_argument("-n", "--name", default=None)
    r.add_argument("-d", "--dir", default=".")
    o = sub.add_parser("org", help="organisation CA signed by the root")
    o.add_argument("--root-cert", required=True)
    o.add_argument("--root-key", required=True)
    o.add_argument("-o", "--organization", required=True)
    o.add_argument("--country", default="US")
    o.add_argument("-n", "--name", default=None)
    o.add_argument("-d", "--dir", default=".")
    n = sub.add_parser("node", help="node certificate (+ chain and trust file) signed by an organisation CA")
    n.add_argument("--ca-cert", required=True)
    n.add_argument("--ca-key", required=True)
    n.add_argument("-n", "--name", required=True)
    n.add_argument("-s", "--san", action="append", default=[])
    n.add_argument("-d", "--dir", default=".")
    c = sub.add_parser("crl", help="revocation list signed by an organisation CA (for the roles' crls_pem)")
    c.add_argument("--ca-cert", required=True)
    c.add_argument("--ca-key", required=True)
    c.add_argument("-r", "--revoke", action="append", default=[], help="node certificate to revoke (repeatable)")
    c.add_argument("-n", "--name", default=None)
    c.add_argument("-d", "--dir", default=".")
    return p


def main(argv=None) -> int:
    a = build_parser().parse_args(argv)
    try:
        if a.command == "root":
            print(*cu.make_root(a.organization, a.name, a.country, a.dir))
        elif a.command == "org":
            print(*cu.make_org(a.root_cert, a.root_key, a.organization, a.name, a.country, a.dir))
        elif a.command == "crl":
            print(cu.make_crl(a.ca_cert, a.ca_key, tuple(a.revoke), a.dir, a.name))
        else:
            print(*cu.make_node(a.ca_cert, a.ca_key, a.name, tuple(a.san), a.dir))
    except (OSError, ValueError) as exc:          # unreadable / malformed CA files, unwritable directory
        print(f"hypha-certutil: error: {exc}", file=sys.stderr)
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
