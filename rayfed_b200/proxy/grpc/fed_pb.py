"""Wire schema of the gRPC transport, built at import time (no protoc needed).

Identical on the wire to the reference's ``fed/grpc/fed.proto:5-19``::

    service GrpcService { rpc SendData (SendDataRequest) returns (SendDataResponse) {} }
    message SendDataRequest  { bytes data = 1; string upstream_seq_id = 2;
                               string downstream_seq_id = 3; string job_name = 4; }
    message SendDataResponse { int32 code = 1; string result = 2; }

The reference ships two generated stub sets picked by protobuf major version
(fed/proxy/grpc/grpc_proxy.py:37-44); building the descriptors programmatically
works for every protobuf ≥ 3.20 and removes generated code from the tree.
``GrpcServiceStub`` / ``GrpcServiceServicer`` /
``add_GrpcServiceServicer_to_server`` mirror what ``fed_pb2_grpc`` exposed.
"""
from __future__ import annotations

import grpc
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

_FILE_NAME = "rayfed_b200/fed.proto"
SERVICE_NAME = "GrpcService"
METHOD_SEND_DATA = f"/{SERVICE_NAME}/SendData"


def _build_file() -> descriptor_pb2.FileDescriptorProto:
    f = descriptor_pb2.FileDescriptorProto()
    f.name = _FILE_NAME
    f.syntax = "proto3"
    T = descriptor_pb2.FieldDescriptorProto

    req = f.message_type.add()
    req.name = "SendDataRequest"
    for number, (name, ftype) in enumerate(
        [("data", T.TYPE_BYTES), ("upstream_seq_id", T.TYPE_STRING), ("downstream_seq_id", T.TYPE_STRING),
         ("job_name", T.TYPE_STRING)],
        start=1,
    ):
        fld = req.field.add()
        fld.name, fld.number, fld.type, fld.label = name, number, ftype, T.LABEL_OPTIONAL

    resp = f.message_type.add()
    resp.name = "SendDataResponse"
    for number, (name, ftype) in enumerate([("code", T.TYPE_INT32), ("result", T.TYPE_STRING)], start=1):
        fld = resp.field.add()
        fld.name, fld.number, fld.type, fld.label = name, number, ftype, T.LABEL_OPTIONAL

    svc = f.service.add()
    svc.name = SERVICE_NAME
    m = svc.method.add()
    m.name = "SendData"
    m.input_type = ".SendDataRequest"
    m.output_type = ".SendDataResponse"
    return f


def _load():
    pool = descriptor_pool.DescriptorPool()
    fd = pool.Add(_build_file()) if hasattr(pool, "Add") else None
    if fd is None:
        pool.AddSerializedFile(_build_file().SerializeToString())
    req_desc = pool.FindMessageTypeByName("SendDataRequest")
    resp_desc = pool.FindMessageTypeByName("SendDataResponse")
    if hasattr(message_factory, "GetMessageClass"):
        return message_factory.GetMessageClass(req_desc), message_factory.GetMessageClass(resp_desc)
    factory = message_factory.MessageFactory(pool)
    return factory.GetPrototype(req_desc), factory.GetPrototype(resp_desc)


SendDataRequest, SendDataResponse = _load()


class GrpcServiceStub:
    """Client stub (sync or ``grpc.aio`` channel)."""

    def __init__(self, channel):
        self.SendData = channel.unary_unary(
            METHOD_SEND_DATA,
            request_serializer=SendDataRequest.SerializeToString,
            response_deserializer=SendDataResponse.FromString,
        )


class GrpcServiceServicer:
    """Base servicer; override ``SendData``."""

    def SendData(self, request, context):
        context.set_code(grpc.StatusCode.UNIMPLEMENTED)
        context.set_details("Method not implemented!")
        raise NotImplementedError("Method not implemented!")


def add_GrpcServiceServicer_to_server(servicer, server):
    handlers = {
        "SendData": grpc.unary_unary_rpc_method_handler(
            servicer.SendData,
            request_deserializer=SendDataRequest.FromString,
            response_serializer=SendDataResponse.SerializeToString,
        )
    }
    server.add_generic_rpc_handlers((grpc.method_handlers_generic_handler(SERVICE_NAME, handlers),))
