"""Self-signed TLS material for tests and local experiments.

Role parity: tool/generate_tls_certs.py:49-125 in the reference (RSA-2048 key +
self-signed certificate written under ``/tmp/rayfed/test-certs/``).  This
version returns the paths, accepts the output directory / subject names as
arguments and adds ``localhost`` + loopback SANs so both gRPC name checks pass.

    python -m rayfed_b200.tool.generate_tls_certs [out_dir]
"""
from __future__ import annotations

import datetime
import ipaddress
import os
import socket
import sys
from typing import Dict, Sequence

DEFAULT_DIR = "/tmp/rayfed/test-certs/"


def generate_self_signed_tls_certs(out_dir: str = DEFAULT_DIR, hostnames: Sequence[str] = ("localhost",),
                                   days_valid: int = 365) -> Dict[str, str]:
    try:
        from cryptography import x509
        from cryptography.hazmat.primitives import hashes, serialization
        from cryptography.hazmat.primitives.asymmetric import rsa
        from cryptography.x509.oid import NameOID
    except ImportError as e:  # pragma: no cover
        raise ImportError("Generating TLS certificates needs the `cryptography` package.") from e

    os.makedirs(os.path.expanduser(out_dir), exist_ok=True)
    try:
        os.chmod(out_dir, 0o777)
    except OSError:
        pass

    key = rsa.generate_private_key(public_exponent=65537, key_size=2048)
    name = x509.Name([x509.NameAttribute(NameOID.COMMON_NAME, "rayfed-b200-test")])
    sans = [x509.DNSName(h) for h in hostnames]
    sans.append(x509.DNSName(socket.gethostname()))
    sans += [x509.IPAddress(ipaddress.ip_address("127.0.0.1")), x509.IPAddress(ipaddress.ip_address("::1"))]
    now = datetime.datetime.now(datetime.timezone.utc)
    cert = (
        x509.CertificateBuilder()
        .subject_name(name)
        .issuer_name(name)
        .public_key(key.public_key())
        .serial_number(x509.random_serial_number())
        .not_valid_before(now - datetime.timedelta(minutes=5))
        .not_valid_after(now + datetime.timedelta(days=days_valid))
        .add_extension(x509.SubjectAlternativeName(sans), critical=False)
        .add_extension(x509.BasicConstraints(ca=True, path_length=None), critical=True)
        .sign(key, hashes.SHA256())
    )
    key_path = os.path.join(out_dir, "server.key")
    cert_path = os.path.join(out_dir, "server.crt")
    ca_path = os.path.join(out_dir, "ca.crt")
    with open(key_path, "wb") as f:
        f.write(key.private_bytes(serialization.Encoding.PEM, serialization.PrivateFormat.TraditionalOpenSSL,
                                  serialization.NoEncryption()))
    pem = cert.public_bytes(serialization.Encoding.PEM)
    for path in (cert_path, ca_path):
        with open(path, "wb") as f:
            f.write(pem)
    return {"ca_cert": ca_path, "cert": cert_path, "key": key_path}


if __name__ == "__main__":
    print(generate_self_signed_tls_certs(sys.argv[1] if len(sys.argv) > 1 else DEFAULT_DIR))
