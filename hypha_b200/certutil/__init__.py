"""Development PKI: Ed25519 root CA -> organisation CA -> node certificate (PKCS#8 keys).

Reference: crates/certutil/src/main.rs:87-372, cli.rs:34-279 (``root -o <org>``,
``org --root-cert --root-key -o <org>``, ``node --ca-cert --ca-key -n <name>``; root valid ten
years starting yesterday).  The loopback fabric of a single box does not need mTLS, but the
tool and the ``cert_pem / key_pem / trust_pem / crls_pem`` config keys are kept so deployments
and config files carry over."""
from __future__ import annotations

import datetime as dt
import os
from typing import Optional, Tuple

from cryptography import x509
from cryptography.hazmat.primitives import serialization
from cryptography.hazmat.primitives.asymmetric.ed25519 import Ed25519PrivateKey
from cryptography.x509.oid import NameOID


def _name(org: str, cn: str, country: str = "US") -> x509.Name:
    return x509.Name([x509.NameAttribute(NameOID.COUNTRY_NAME, country),
                      x509.NameAttribute(NameOID.ORGANIZATION_NAME, org),
                      x509.NameAttribute(NameOID.COMMON_NAME, cn)])


def _key_pem(key: Ed25519PrivateKey) -> bytes:
    return key.private_bytes(serialization.Encoding.PEM, serialization.PrivateFormat.PKCS8, serialization.NoEncryption())


def _write(path: str, data: bytes, mode: int) -> None:
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)       # `-d pki` may not exist yet
    fd = os.open(path, os.O_WRONLY | os.O_CREAT | os.O_TRUNC, mode)
    with os.fdopen(fd, "wb") as f:
        f.write(data)


def _build(subject: x509.Name, issuer: x509.Name, pub, signer: Ed25519PrivateKey, days: int, ca: bool,
           path_len: Optional[int], sans: Tuple[str, ...] = ()) -> x509.Certificate:
    now = dt.datetime.now(dt.timezone.utc)
    b = (x509.CertificateBuilder().subject_name(subject).issuer_name(issuer).public_key(pub)
         .serial_number(x509.random_serial_number()).not_valid_before(now - dt.timedelta(days=1))
         .not_valid_after(now + dt.timedelta(days=days))
         .add_extension(x509.BasicConstraints(ca=ca, path_length=path_len), critical=True))
    if ca:
        b = b.add_extension(x509.KeyUsage(digital_signature=True, key_cert_sign=True, crl_sign=True,
                                          content_commitment=False, key_encipherment=False, data_encipherment=False,
                                          key_agreement=False, encipher_only=False, decipher_only=False), critical=True)
    else:
        b = b.add_extension(x509.ExtendedKeyUsage([x509.oid.ExtendedKeyUsageOID.SERVER_AUTH,
                                                   x509.oid.ExtendedKeyUsageOID.CLIENT_AUTH]), critical=False)
        if sans:
            b = b.add_extension(x509.SubjectAlternativeName([x509.DNSName(s) for s in sans]), critical=False)
    return b.sign(signer, algorithm=None)


def make_root(org: str, name: Optional[str] = None, country: str = "US", out_dir: str = ".") -> Tuple[str, str]:
    name = name or f"{org}-root"
    key = Ed25519PrivateKey.generate()
    subj = _name(org, f"{org} Root CA", country)
    cert = _build(subj, subj, key.public_key(), key, 3650, True, 1)
    cp, kp = os.path.join(out_dir, f"{name}-cert.pem"), os.path.join(out_dir, f"{name}-key.pem")
    _write(cp, cert.public_bytes(serialization.Encoding.PEM), 0o644)
    _write(kp, _key_pem(key), 0o600)
    return cp, kp


def _load(cert_path: str, key_path: str):
    cert = x509.load_pem_x509_certificate(open(cert_path, "rb").read())
    key = serialization.load_pem_private_key(open(key_path, "rb").read(), password=None)
    return cert, key


def make_org(root_cert: str, root_key: str, org: str, name: Optional[str] = None, country: str = "US",
             out_dir: str = ".") -> Tuple[str, str]:
    name = name or f"{org}-ca"
    rc, rk = _load(root_cert, root_key)
    key = Ed25519PrivateKey.generate()
    cert = _build(_name(org, f"{org} CA", country), rc.subject, key.public_key(), rk, 1825, True, 0)
    cp, kp = os.path.join(out_dir, f"{name}-cert.pem"), os.path.join(out_dir, f"{name}-key.pem")
    _write(cp, cert.public_bytes(serialization.Encoding.PEM) + rc.public_bytes(serialization.Encoding.PEM), 0o644)
    _write(kp, _key_pem(key), 0o600)
    return cp, kp


def make_node(ca_cert: str, ca_key: str, name: str, sans: Tuple[str, ...] = (), out_dir: str = ".") -> Tuple[str, str, str]:
    chain = open(ca_cert, "rb").read()
    cc, ck = _load(ca_cert, ca_key)
    org = cc.subject.get_attributes_for_oid(NameOID.ORGANIZATION_NAME)[0].value
    key = Ed25519PrivateKey.generate()
    cert = _build(_name(org, name), cc.subject, key.public_key(), ck, 365, False, None, sans or (name,))
    cp, kp, tp = (os.path.join(out_dir, f"{name}-{s}.pem") for s in ("cert", "key", "trust"))
    _write(cp, cert.public_bytes(serialization.Encoding.PEM) + chain, 0o644)
    _write(kp, _key_pem(key), 0o600)
    certs = x509.load_pem_x509_certificates(chain)
    _write(tp, certs[-1].public_bytes(serialization.Encoding.PEM), 0o644)
    return cp, kp, tp


def make_crl(ca_cert: str, ca_key: str, revoked: Tuple[str, ...] = (), out_dir: str = ".", name: Optional[str] = None,
             days: int = 30) -> str:
    """Certificate revocation list signed by the organisation CA for the node certificates in ``revoked`` (paths
    of ``<name>-cert.pem`` files) -- what the roles' ``crls_pem`` key points at (the reference loads CRLs,
    crates/network/src/cert.rs, but ships no tool to make one)."""
    cc, ck = _load(ca_cert, ca_key)
    now = dt.datetime.now(dt.timezone.utc)
    b = (x509.CertificateRevocationListBuilder().issuer_name(cc.subject)
         .last_update(now - dt.timedelta(minutes=5)).next_update(now + dt.timedelta(days=days)))
    for path in revoked:
        leaf = x509.load_pem_x509_certificates(open(path, "rb").read())[0]
        b = b.add_revoked_certificate(x509.RevokedCertificateBuilder().serial_number(leaf.serial_number)
                                      .revocation_date(now - dt.timedelta(minutes=5)).build())
    crl = b.sign(ck, None)
    out = os.path.join(out_dir, f"{name or 'revoked'}-crl.pem")
    _write(out, crl.public_bytes(serialization.Encoding.PEM), 0o644)
    return out


def verify_chain(node_cert_path: str) -> bool:
    """Signature check of every link of the chain stored in ``<name>-cert.pem``."""
    certs = x509.load_pem_x509_certificates(open(node_cert_path, "rb").read())
    for child, parent in zip(certs, certs[1:]):
        parent.public_key().verify(child.signature, child.tbs_certificate_bytes)
    certs[-1].public_key().verify(certs[-1].signature, certs[-1].tbs_certificate_bytes)
    return True
